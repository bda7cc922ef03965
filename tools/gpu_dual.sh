#!/bin/bash
set +e
O=gpurun_out; TAG=${1:-dual}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/${TAG}_tests.txt
for v in "ELECTOR_NO_DUAL=1" "ELECTOR_PH2D_WARPS=24" "ELECTOR_PH2D_WARPS=20"; do
  echo "== $v" | tee -a $O/${TAG}_step.txt
  env $v ELECTOR_TRACE=2 python tools/seg_trace.py 10000 1 1 2>&1 | grep -E "^call 2|phase 2 segment +(9|10|11) |all done" | tail -5 | tee -a $O/${TAG}_step.txt
done
