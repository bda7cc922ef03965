#!/bin/bash
# GPU check of the worker pipeline: parity tests, then the host-buffer call at several chunk / worker counts
set +e
O=gpurun_out; TAG=${1:-pipe}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
for w in 1 2 3 4; do
  echo "== workers $w" | tee -a $O/${TAG}_pipe.txt
  ELECTOR_PIPELINE_WORKERS=$w DIAG_CHUNKS=${2:-1,3,6,8,12,16} timeout 400 python tools/pipe_diag.py 10000 1 2>&1 | grep -v "trace\|pinned\|together" | tee -a $O/${TAG}_pipe.txt
done
echo "== all segments on side streams, workers 3" | tee -a $O/${TAG}_pipe.txt
ELECTOR_ALL_SIDE=1 ELECTOR_PIPELINE_WORKERS=3 DIAG_CHUNKS=1,8 timeout 400 python tools/pipe_diag.py 10000 1 2>&1 | grep -v "trace\|pinned\|together" | tee -a $O/${TAG}_pipe.txt
