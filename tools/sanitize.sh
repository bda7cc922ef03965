#!/bin/bash
# compute-sanitizer evidence: memcheck and racecheck of the whole pipelined call (POA both phases,
# merge, tally) through the C driver (no Python in the checked process), on slices of several configs, and of the `poa`
# drop-in on the golden sets (edge cases, long windows).
#   bash tools/sanitize.sh TAG        -> gpurun_out/TAG_sanitizer.txt
set +e
O=gpurun_out; TAG=${1:-r2}; mkdir -p $O; L=$O/${TAG}_sanitizer.txt; : > $L
python -c "import elector_b200; elector_b200.write_default_matrix('/tmp/blosum80.mat')"
for spec in "4 240" "2 60" "1 40"; do
  set -- $spec
  python tools/dump_csr.py $2 $1 /tmp/san_c$1 > /dev/null
  for tool in memcheck racecheck; do
    echo "== $tool: pipe_driver, config $1, $2 reads" >> $L
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 elector_b200/bin/pipe_driver /tmp/san_c$1 2 2>&1 | grep -v "Host Frame\|Saved host backtrace\|^=========$" | tail -25 >> $L
  done
done
for set in edge hard long; do
  for e in ref.fa cor.fa unc.fa; do zcat tests/golden/$set.$e.gz > /tmp/san_$set.$e; done
  echo "== memcheck: poa drop-in, golden set $set" >> $L
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 elector_b200/bin/poa -pir /tmp/san_$set.pir -corrected_reads_fasta /tmp/san_$set.cor.fa -reference_reads_fasta /tmp/san_$set.ref.fa -uncorrected_reads_fasta /tmp/san_$set.unc.fa -pathMatrix /tmp/blosum80.mat 2>&1 | grep -v "Host Frame\|Saved host backtrace\|^=========$\|^0 1 2" | tail -12 >> $L
  zcat tests/golden/$set.pir.gz | cmp - /tmp/san_$set.pir >> $L 2>&1 && echo "   PIR identical to the reference under the sanitizer" >> $L
done
grep -c "ERROR SUMMARY: 0 errors" $L
