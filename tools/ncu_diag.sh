#!/bin/bash
# full ncu capture of the DP kernels through the C `poa` shim (no Python in the profiled process:
# ncu's multi-pass replay segfaults inside the interpreter on this image)
set +e
O=gpurun_out
TAG=${1:-r1c}
READS=${2:-2000}
mkdir -p $O
exec > $O/${TAG}_ncu.log 2>&1
python tools/dump_fasta.py $READS 1 /tmp/prof
python -c "import elector_b200; elector_b200.write_default_matrix('/tmp/blosum80.mat')"
CMD="elector_b200/bin/poa -pir /tmp/prof.pir -corrected_reads_fasta /tmp/prof.cor.fa -reference_reads_fasta /tmp/prof.ref.fa -uncorrected_reads_fasta /tmp/prof.unc.fa -pathMatrix /tmp/blosum80.mat"
$CMD > /dev/null; echo plain rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_dp -c 10 -o $O/${TAG}_full -f $CMD > /dev/null; echo rc=$?
ls -la $O
