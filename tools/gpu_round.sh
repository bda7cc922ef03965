#!/bin/bash
# One GPU-box round: parity tests, bench line, ncu launch list, ncu full capture of the DP kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
# Everything lands under gpurun_out/TAG_*; copy what should be judged into profiles/.
set +e
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 3000 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python tools/profile_step.py 10000 2 1 pipeline > $O/${TAG}_launches.log 2>&1; tail -2 $O/${TAG}_launches.log
echo "== ncu full (bulk DP launches, through the C poa shim)"
bash tools/ncu_diag.sh ${TAG} 2000; tail -3 $O/${TAG}_ncu.log
ls -la $O | tail -12
