#!/bin/bash
# quick GPU check: parity tests, then kernel time of config 1 through the host C-ABI (packed and INT32 kernels)
set +e
O=gpurun_out; TAG=${1:-q}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
timeout 600 python tools/profile_step.py 10000 4 1 poa 2>&1 | tail -4 | tee $O/${TAG}_step.txt
ELECTOR_NO_PACKED=1 timeout 600 python tools/profile_step.py 10000 3 1 poa 2>&1 | tail -2 | tee $O/${TAG}_step_int32.txt
