#!/bin/bash
# resident-warp caps (scratch footprint in L2 vs latency hiding): kernel time of one config-1 call per setting
set +e
O=gpurun_out; TAG=${1:-caps}; mkdir -p $O
for v in "X=0" "ELECTOR_WARPS_PH2L=24" "ELECTOR_WARPS_PH2L=16" "ELECTOR_WARPS_PH1P=24" "ELECTOR_WARPS_PH1P=16" "ELECTOR_WARPS_PH2D=16" "ELECTOR_WARPS_PH2D=12" "ELECTOR_WARPS_PH2L=20 ELECTOR_WARPS_PH1P=20 ELECTOR_WARPS_PH2D=16"; do
  echo "== $v" | tee -a $O/${TAG}.txt
  env $v python tools/seg_trace.py 10000 1 1 2>&1 | grep -E "^call 2" | tee -a $O/${TAG}.txt
done
