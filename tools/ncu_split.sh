#!/bin/bash
# Source-level ncu capture of the window-cutting kernel (split_jobs_kernel) through the masterSplitter drop-in executable:
#   ${TAG}_split_summary.csv, ${TAG}_split_stalls.txt, ${TAG}_split_byfunc.txt (per function / line: instruction and stall-sample shares)
set +e
O=gpurun_out; TAG=${1:-r2p}; CFG=${2:-1}; READS=${3:-2000}; mkdir -p $O /tmp/sp_out
python -c "import workloads; print(workloads.ensure_gen())" > /dev/null
tools/gen_reads $CFG $READS 0 /tmp/sp$READS
A="/tmp/sp$READS.ref.fa /tmp/sp$READS.unc.fa /tmp/sp$READS.cor.fa /tmp/sp_out/out1 /tmp/sp_out/out2 /tmp/sp_out/out3 7 200 10000 0.1 /tmp/sp_out"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:split_jobs_kernel -c 1 -o /tmp/${TAG}_split -f elector_b200/bin/masterSplitter $A > /dev/null
echo "capture rc=$?"
mkdir -p /tmp/dis && (cd /tmp/dis && cuobjdump -xelf all $OLDPWD/elector_b200/libelector_poa.so > /dev/null && nvdisasm -g -c capi.sm_100a.cubin > dis.txt 2>/dev/null)
python tools/ncu_summary.py /tmp/${TAG}_split.ncu-rep > $O/${TAG}_split_summary.csv
python tools/ncu_stalls.py /tmp/${TAG}_split.ncu-rep > $O/${TAG}_split_stalls.txt
ncu -i /tmp/${TAG}_split.ncu-rep --page source --csv > /tmp/split_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/split_src.csv /tmp/dis/dis.txt split_jobs_kernel 1000 > /tmp/split_lines.txt
f=$O/${TAG}_split_byfunc.txt
python tools/ncu_funcs.py < /tmp/split_lines.txt > $f
echo "# ---- lines" >> $f
head -80 /tmp/split_lines.txt >> $f
cp /tmp/${TAG}_split.ncu-rep $O/ 2>/dev/null
