#!/bin/bash
set +e
O=gpurun_out; TAG=${1:-coop2}; mkdir -p $O
export ELECTOR_PIPELINE_CHUNKS=1
for g in 1 2 4 8; do
  ELECTOR_COOP_GROUP=$g timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launch_g$g.csv python tools/profile_step.py 10000 2 1 poa > $O/${TAG}_g$g.log 2>&1
done
for g in 0 1 2 4; do
  echo "== 1 chunk, ELECTOR_COOP_GROUP=$g" | tee -a $O/${TAG}_step.txt
  ELECTOR_COOP_GROUP=$g timeout 600 python tools/profile_step.py 10000 4 1 poa 2>&1 | tail -2 | tee -a $O/${TAG}_step.txt
done
