#!/bin/bash
# One GPU-box round: parity tests, smoke, bench line (+ reference arm), ncu evidence through the C `poa` executable.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh TAG'
# Everything lands under gpurun_out/TAG_*; copy what should be judged into profiles/.
set +e
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/${TAG}_tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 1500 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench_reference.json
echo "== ncu"; bash tools/ncu_round.sh ${TAG}; tail -4 $O/${TAG}_ncu.log
