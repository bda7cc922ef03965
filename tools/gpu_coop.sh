#!/bin/bash
# GPU check of the warp-cooperative kernel: parity tests, then kernel time of config 1 with the longest windows
# thread-per-window (group 0) and warp-cooperative at several group sizes, then the host-buffer pipeline timeline
set +e
O=gpurun_out; TAG=${1:-coop}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
for g in 0 4 8 16; do
  echo "== ELECTOR_COOP_GROUP=$g" | tee -a $O/${TAG}_step.txt
  ELECTOR_COOP_GROUP=$g timeout 600 python tools/profile_step.py 10000 4 1 poa 2>&1 | tail -2 | tee -a $O/${TAG}_step.txt
done
DIAG_CHUNKS=1,2,3,4 timeout 400 python tools/pipe_diag.py 10000 1 > $O/${TAG}_pipe.txt 2>&1; grep -v trace $O/${TAG}_pipe.txt
