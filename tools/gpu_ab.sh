#!/bin/bash
# A/B on the GPU box: parity tests, then kernel time of config 1 through the host C-ABI with experiment knobs
#   gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh TAG'
set +e
O=gpurun_out; TAG=${1:-ab}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
run() { echo "== $*"; env "$@" timeout 600 python tools/profile_step.py 10000 4 1 poa 2>&1 | tail -2; }
{
run A=default
run ELECTOR_NO_PACKED2=1
run ELECTOR_WARPS_PH2P=20
run ELECTOR_WARPS_PH2P=16
run ELECTOR_WARPS_PH2P=12
run ELECTOR_WARPS_PH1P=24
run ELECTOR_WARPS_PH1P=16
} 2>&1 | tee $O/${TAG}_ab.txt
echo "== pipeline trace"
ELECTOR_TRACE=1 timeout 600 python tools/profile_step.py 10000 3 1 pipeline 2>&1 | tail -14 | tee $O/${TAG}_trace.txt
