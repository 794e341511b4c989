REPS=2 timeout 300 bash tools/abn.sh "c2 c1" speechain_b200/csrc/libsfe.so tools/ab/libsfe_w20.so tools/ab/libsfe_w16.so
