#!/bin/bash
# Source-level ncu capture of the bulk DP launches of one 10 000-read call (one chunk, through the C `poa` executable),
# attributed on the box to source lines and functions (tools/ncu_lines.py, tools/ncu_funcs.py):
#   ${TAG}_byfunc_<k>_<kernel>.txt : per function and per line share of executed warp instructions and of stall samples
#   ${TAG}_src_summary.csv / ${TAG}_src_stalls.txt : per-launch summary and stall reasons of the same capture
set +e
O=gpurun_out; TAG=${1:-r2a}; FILTER=${2:-regex:poa_dp[12]_kernel}; COUNT=${3:-9}; READS=${4:-10000}; mkdir -p $O
export ELECTOR_PIPELINE_CHUNKS=1 ELECTOR_PIPELINE_WORKERS=1 ELECTOR_SERVICE=0   # (the executables run in process: ncu follows no server)
python -c "import elector_b200; elector_b200.write_default_matrix('/tmp/blosum80.mat')"
[ -f /tmp/prof$READS.ref.fa ] || python tools/dump_fasta.py $READS 1 /tmp/prof$READS
timeout 1200 ncu --set full --clock-control none --import-source on -k $FILTER -c $COUNT -o /tmp/${TAG}_src -f elector_b200/bin/poa -pir /tmp/prof$READS.pir -corrected_reads_fasta /tmp/prof$READS.cor.fa -reference_reads_fasta /tmp/prof$READS.ref.fa -uncorrected_reads_fasta /tmp/prof$READS.unc.fa -pathMatrix /tmp/blosum80.mat > /dev/null
echo "byfunc capture rc=$?"
mkdir -p /tmp/dis && (cd /tmp/dis && cuobjdump -xelf all $OLDPWD/elector_b200/libelector_poa.so > /dev/null && nvdisasm -g -c capi.sm_100a.cubin > dis.txt 2>/dev/null)
python tools/ncu_summary.py /tmp/${TAG}_src.ncu-rep > $O/${TAG}_src_summary.csv
python tools/ncu_stalls.py /tmp/${TAG}_src.ncu-rep > $O/${TAG}_src_stalls.txt
python - $O/${TAG}_src_summary.csv <<'PY' > /tmp/kernels.txt
import csv, sys
for k, r in enumerate(list(csv.reader(open(sys.argv[1])))[1:]):
    print(k, r[0].replace(" ", ""), r[1].replace(" ", ""))
PY
while read k name grid; do
  ncu -i /tmp/${TAG}_src.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > /tmp/src_$k.csv 2>/dev/null
  # the mangled name holds the template argument: Phase2L -> 7Phase2LE
  key=$(echo "$name" | sed -E 's/.*<(elector::)?([A-Za-z0-9]+).*/\2/')
  f=$O/${TAG}_byfunc_${k}_${key}.txt
  echo "# $name grid $grid" > $f
  python tools/ncu_lines.py /tmp/src_$k.csv /tmp/dis/dis.txt "${key}E" 1000 > /tmp/lines_$k.txt
  python tools/ncu_funcs.py < /tmp/lines_$k.txt >> $f
  echo "# ---- lines" >> $f
  head -120 /tmp/lines_$k.txt >> $f
done < /tmp/kernels.txt
ls -la $O | head -40
