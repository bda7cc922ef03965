#!/bin/bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference binaries for the
# POA hot path from the sources where they lie under /root/reference, with
# outputs only into oracle/_ref/ (git-ignored; travels to the GPU box).
# No reference source is copied into this repository.
#
#   oracle/_ref/poa             reference src/poa-graph (main.c + liblpo objects)
#   oracle/_ref/masterSplitter  reference src/split/Master_Splitter.cpp
#   oracle/_ref/Donatello       reference src/split/Donatello.cpp
#   oracle/_ref/ref_dump        our harness (oracle/ref_harness.c) linked against
#                               the reference's own objects: dumps DP scores and
#                               alignments that the poa binary never prints
#   oracle/_ref/blosum80.mat    the matrix file (data, 42 lines) next to the binaries
#   oracle/_ref/elector_tree/   an ELECTOR installation as install.sh lays it out, for running the reference's OWN Python
#                               (alignment.getPOA, computeStats.outputRecallPrecision) next to the replacements on the GPU box:
#                               elector/*.py (unmodified copies), bin/{poa,masterSplitter,Donatello} (the builds above),
#                               src/poa-graph/blosum80.mat, and an empty Bio/ stub (the modules import Bio.SeqIO at the top
#                               but the two functions never use it; Biopython is not in the image)
#
# Flag-level workarounds only: -fcommon (GCC>=10 vs
# black_flag.h tentative definitions), -include cstdint (Master_Splitter.cpp).
set -euo pipefail
REF=${ELECTOR_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/poa-graph" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
P="$REF/src/poa-graph"
CF="-g -Ofast -DUSE_WEIGHTED_LINKS -DUSE_PROJECT_HEADER -I$P -pthread -fcommon -w"
LIB="black_flag seq_util fasta_format msa_format align_lpo2 align_lpo_po2 buildup_lpo lpo heaviest_bundle lpo_format create_seq remove_bundle numeric_data stringptr"
OBJS=""
for f in $LIB align_score main; do
  gcc $CF -c "$P/$f.c" -o "$OUT/obj/$f.o" &
done
wait
for f in $LIB; do OBJS="$OBJS $OUT/obj/$f.o"; done
gcc -o "$OUT/poa" "$OUT/obj/align_score.o" "$OUT/obj/main.o" $OBJS -lm -lpthread
gcc $CF -c "$HERE/ref_harness.c" -o "$OUT/obj/ref_harness.o"
gcc -o "$OUT/ref_dump" "$OUT/obj/ref_harness.o" "$OUT/obj/align_score.o" $OBJS -lm -lpthread
g++ -w -Ofast -std=c++11 -fopenmp -include cstdint "$REF/src/split/Master_Splitter.cpp" -o "$OUT/masterSplitter"
g++ -w -Ofast -std=c++11 "$REF/src/split/Donatello.cpp" -o "$OUT/Donatello"
cp "$P/blosum80.mat" "$OUT/blosum80.mat"
T="$OUT/elector_tree"
rm -rf "$T"
mkdir -p "$T/elector" "$T/bin" "$T/src/poa-graph" "$T/Bio"
cp "$REF"/elector/*.py "$T/elector/"
cp "$OUT/poa" "$OUT/masterSplitter" "$OUT/Donatello" "$T/bin/"
cp "$P/blosum80.mat" "$T/src/poa-graph/blosum80.mat"
printf '' > "$T/Bio/__init__.py"
printf '' > "$T/Bio/SeqIO.py"
printf 'class Seq(str):\n    pass\n' > "$T/Bio/Seq.py"
rm -rf "$OUT/obj"
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
