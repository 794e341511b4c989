REPS=2 timeout 400 bash tools/abn.sh "c2 c1" speechain_b200/csrc/libsfe.so tools/ab/libsfe_w22.so tools/ab/libsfe_mu2.so tools/ab/libsfe_mu3.so
