timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
REPS=1 timeout 300 bash tools/abn.sh "c5 c3 c2" tools/ab/libsfe_prev.so speechain_b200/csrc/libsfe.so
