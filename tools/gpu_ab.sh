#!/bin/bash
# A/B on the GPU box: parity tests, then kernel time of config 1 through the host C-ABI with experiment knobs
#   gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh TAG'
set +e
O=gpurun_out; TAG=${1:-ab}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
run() { echo "== $*"; env "$@" timeout 600 python tools/profile_step.py 10000 4 1 poa 2>&1 | tail -2; }
{
run A=default
run ELECTOR_NO_LINEAR2=1
run ELECTOR_PACKED2=1
run ELECTOR_WARPS_PH2L=24
run ELECTOR_WARPS_PH2L=16
} 2>&1 | tee $O/${TAG}_ab.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 2500 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
echo "== pipeline diag (pinned)"
DIAG_CHUNKS=1,2,3,4,6 timeout 600 python tools/pipe_diag.py 10000 1 2>&1 | tail -40 | tee $O/${TAG}_pipe.txt
