#!/bin/bash
O=gpurun_out/r2s18; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest.log; cat $O/pytest.log
