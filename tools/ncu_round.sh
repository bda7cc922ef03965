#!/bin/bash
# ncu evidence of one round through the C `poa` executable (no Python in the profiled process):
#   ${TAG}_launches.csv : every kernel launch of one 10 000-read call with its duration (cold-cache, serialised)
#   ${TAG}_full.ncu-rep : --set full of the first POA launches of a 2 000-read call (bulk launches included)
#   ${TAG}_launches_pipeline.csv : launch list of two pipelined calls (POA + merge + tally) through tools/pipe_driver.c
#   ${TAG}_pipe_driver.txt / ${TAG}_pipe_trace.txt : host-clock time of 8 calls; device timeline (ELECTOR_TRACE=2) of one
#   ${TAG}_launches_metrics.csv / ${TAG}_issue.json / ${TAG}_launch_table.csv : the pipelined call's launches with instruction counts, IPC, ALU-pipe use, DRAM bytes
#   ${TAG}_dp2_10k.ncu-rep + ${TAG}_traffic.json : --set full of the phase-2 launch set of the 10 000-read call (DRAM bytes per launch)
set +e
O=gpurun_out; TAG=${1:-r1d}; mkdir -p $O
exec > $O/${TAG}_ncu.log 2>&1
# the profiled calls run in ONE chunk on one worker, like the device-resident step bench.py's `value` and roofline time
export ELECTOR_PIPELINE_CHUNKS=1 ELECTOR_PIPELINE_WORKERS=1 ELECTOR_SERVICE=0   # (the executables run in process: ncu follows no server)
python -c "import elector_b200; elector_b200.write_default_matrix('/tmp/blosum80.mat')"
for R in 10000 2000; do python tools/dump_fasta.py $R 1 /tmp/prof$R; done
cmd() { echo "elector_b200/bin/poa -pir /tmp/prof$1.pir -corrected_reads_fasta /tmp/prof$1.cor.fa -reference_reads_fasta /tmp/prof$1.ref.fa -uncorrected_reads_fasta /tmp/prof$1.unc.fa -pathMatrix /tmp/blosum80.mat"; }
$(cmd 2000) > /dev/null; echo plain rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches.csv $(cmd 10000) > /dev/null; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_dp -c 14 -o $O/${TAG}_full -f $(cmd 2000) > /dev/null; echo full rc=$?
# the whole pipelined call (POA + merge + tally) in one chunk on one worker, through the C driver of elector_pipeline_run
python tools/dump_csr.py 10000 1 /tmp/c1 > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_pipeline.csv elector_b200/bin/pipe_driver /tmp/c1 2 > /dev/null; echo pipeline launches rc=$?
# the same two calls with instruction counts, IPC and ALU-pipe use per launch -> ${TAG}_issue.json (bench.py: roofline.issue)
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_metrics.csv elector_b200/bin/pipe_driver /tmp/c1 2 > /dev/null; echo launch metrics rc=$?
python tools/ncu_issue.py $O/${TAG}_launches_metrics.csv $O/${TAG}_issue.json
python tools/launch_metrics.py $O/${TAG}_launches_metrics.csv "poa_dp|merge|tally|bin" > $O/${TAG}_launch_table.csv
# --set full of the phase-2 launch set of the SECOND call (the first one grows the scratch pools and runs segments twice)
set -- $(python tools/ncu_skip.py $O/${TAG}_launches_pipeline.csv poa_dp2); echo "phase-2 launches: skip $1, capture $2"
# (every poa_dp2 launch of the run is captured and the last $2 are kept: under the full set the first call's launch pattern differs)
timeout 1500 ncu --set full --clock-control none -k regex:poa_dp2 -c 80 -o $O/${TAG}_dp2_10k -f elector_b200/bin/pipe_driver /tmp/c1 3 > /dev/null; echo dp2-10k rc=$?
python tools/ncu_traffic.py $O/${TAG}_dp2_10k.ncu-rep $O/${TAG}_traffic.json $2
# summaries are made here; the reports themselves (30 MB each) stay on the box unless KEEP_REPS=1 (gpurun_out/ is capped at 64 MiB)
python tools/ncu_summary.py $O/${TAG}_full.ncu-rep > $O/${TAG}_ncu_full_summary_2k_reads.csv
python tools/ncu_summary.py $O/${TAG}_dp2_10k.ncu-rep > $O/${TAG}_ncu_full_summary_dp2_10k_reads.csv
python tools/ncu_stalls.py $O/${TAG}_dp2_10k.ncu-rep > $O/${TAG}_ncu_stalls_dp2_10k_reads.txt
[ "$KEEP_REPS" = 1 ] || rm -f $O/${TAG}_full.ncu-rep $O/${TAG}_dp2_10k.ncu-rep
unset ELECTOR_PIPELINE_CHUNKS ELECTOR_PIPELINE_WORKERS   # the default pipeline (chunks on worker contexts) for the host-clock numbers
elector_b200/bin/pipe_driver /tmp/c1 8 > $O/${TAG}_pipe_driver.txt 2>&1
ELECTOR_TRACE=2 elector_b200/bin/pipe_driver /tmp/c1 3 2>&1 | tail -60 > $O/${TAG}_pipe_trace.txt
ls -la $O
