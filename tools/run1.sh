timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
REPS=1 timeout 300 bash tools/abn.sh "c2 c1 c3" tools/ab/libsfe_base.so tools/ab/libsfe_melv2.so tools/ab/libsfe_atom.so speechain_b200/csrc/libsfe.so
