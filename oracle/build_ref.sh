#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference `biokanga` executable from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU
# box with the snapshot).  No reference source file is copied into this repository: every
# translation unit is compiled in place, objects and the binary land in oracle/_ref/.
#
# One translation unit (biokanga/Aligner.cpp) is compiled from a sed-patched stream on stdin:
# the reference hands a stack-local tsLoadReadsThreadPars to its loader thread
# (Aligner.cpp:4822,4838) and returns after <=3 s (Aligner.cpp:4853); the loader later writes
# through the dangling pointer (Aligner.cpp:4810).  Making that one local `static` fixes the
# use-after-return and does not change any result.
#
# The reference's own build system (autotools) is NOT run; this is the hand g++ recipe of
# SURVEY.md section 8(c).  libbiokanga/sqlite3.c is absent from the checkout, the system
# libsqlite3.so.0 is linked instead; system -lz replaces the vendored zlib.
set -euo pipefail
REF=${BKX_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/biokanga" ]; then
  echo "build_ref: $REF not present (GPU box?) - using prebuilt $OUT/biokanga if any" >&2
  exit 0
fi
if [ -x "$OUT/biokanga" ] && [ "$OUT/biokanga" -nt "$0" ]; then
  echo "build_ref: $OUT/biokanga up to date"; exit 0
fi
mkdir -p "$OUT/obj/lib" "$OUT/obj/bk" "$OUT/obj/pl"
CXXFLAGS="-O2 -w -fpermissive -std=gnu++11 -D_LARGEFILE64_SOURCE -D_FILE_OFFSET_BITS=64"
SQLITE=$(ls /lib/x86_64-linux-gnu/libsqlite3.so.0 /usr/lib/x86_64-linux-gnu/libsqlite3.so.0 2>/dev/null | head -1)

LIBSRC="AlignValidate argtable2 BEDfile BioSeqFile Centroid Conformation ConfSW CSVFile CVS2BED DataPoints
 Diagnostics Endian ErrorCodes Fasta FeatLoci FilterLoci FilterRefIDs GOAssocs GOTerms HashFile HyperEls
 GFFFile GTFFile Contaminants MAlignFile Random SimpleRNG RsltsFile sais SAMfile SeqTrans SfxArray SfxArrayV2
 Shuffle SmithWaterman NeedlemanWunsch Stats StopWatch Twister Utility ProcRawReads MTqsort bgzf"

cmds=$OUT/obj/cmds.txt; : > "$cmds"
for s in $LIBSRC; do
  echo "cd $REF/libbiokanga && g++ $CXXFLAGS -c $s.cpp -o $OUT/obj/lib/$s.o" >> "$cmds"
done
for f in "$REF"/biokanga/*.cpp; do
  b=$(basename "$f" .cpp)
  [ "$b" = stdafx ] && continue
  if [ "$b" = Aligner ]; then
    echo "cd $REF/biokanga && sed '4822s/^tsLoadReadsThreadPars ThreadPars;/static tsLoadReadsThreadPars ThreadPars;/' Aligner.cpp | g++ $CXXFLAGS -x c++ -I$REF/biokanga -c - -o $OUT/obj/bk/$b.o" >> "$cmds"
  else
    echo "cd $REF/biokanga && g++ $CXXFLAGS -c $b.cpp -o $OUT/obj/bk/$b.o" >> "$cmds"
  fi
done
for f in "$REF"/libBKPLPlot/*.cpp; do
  b=$(basename "$f" .cpp)
  [ "$b" = stdafx ] && continue
  extra=""
  [ "$b" = plstdio ] && extra="-DO_BINARY=0 -D_O_SHORT_LIVED=0 -D_O_TEMPORARY=0"
  echo "cd $REF/libBKPLPlot && g++ $CXXFLAGS $extra -c $b.cpp -o $OUT/obj/pl/$b.o" >> "$cmds"
done
xargs -d "\n" -P "$JOBS" -I{} bash -c "{}" < "$cmds"
rm -f "$OUT/obj/libbiokanga.a" "$OUT/obj/libBKPLPlot.a"
ar rcs "$OUT/obj/libbiokanga.a" "$OUT"/obj/lib/*.o
ar rcs "$OUT/obj/libBKPLPlot.a" "$OUT"/obj/pl/*.o
g++ -O2 -o "$OUT/biokanga" "$OUT"/obj/bk/*.o "$OUT/obj/libbiokanga.a" "$OUT/obj/libBKPLPlot.a" \
    -lz "$SQLITE" -lpthread -ldl -lrt
rm -rf "$OUT/obj"
echo "build_ref: built $OUT/biokanga"
