#!/bin/bash
O=gpurun_out/r2s37; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
for lib in tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so tools/ab/libsfe_padinl.so; do echo "== $lib"; SFE_LIB_PATH=$PWD/$lib python tools/c2gap.py 2>&1 | grep "ragged"; done | tee $O/c2gap.txt
REPS=2 bash tools/abn.sh "c1 c2" tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so tools/ab/libsfe_padinl.so 2>&1 | tee $O/abn.log
for lib in speechain_b200/csrc/libsfe.so tools/ab/libsfe_p11nf16.so tools/ab/libsfe_p11nf12.so tools/ab/libsfe_p11nf10.so; do echo "== $lib"; SFE_LIB_PATH=$PWD/$lib python tools/cfg400.py 2>&1 | tail -4; done | tee $O/cfg400.txt
