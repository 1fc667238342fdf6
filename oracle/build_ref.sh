#!/bin/sh
# Build the UNMODIFIED reference C library (sources read in place from
# $TSKIT_REFERENCE, default /root/reference) plus oracle/ref_shim.c into
# oracle/_ref/libtskit_ref.so.  Test infrastructure only.  No reference source
# is copied; only the built .so lands in the (git-ignored) oracle/_ref/.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TSKIT_REFERENCE:-/root/reference}"
if [ ! -d "$REF/c/tskit" ]; then
    echo "build_ref.sh: $REF/c/tskit not found; keeping any prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$HERE/_ref"
OUT="$HERE/_ref/libtskit_ref.so"
SRCS="$REF/c/tskit/core.c $REF/c/tskit/tables.c $REF/c/tskit/trees.c \
 $REF/c/tskit/genotypes.c $REF/c/tskit/stats.c $REF/c/tskit/convert.c \
 $REF/c/tskit/haplotype_matching.c $REF/c/subprojects/kastore/kastore.c"
NEWER=0
if [ ! -f "$OUT" ] || [ "$HERE/ref_shim.c" -nt "$OUT" ]; then NEWER=1; fi
if [ "$NEWER" = 1 ]; then
    # same optimisation level and dialect as the reference's python/setup.py
    gcc -O2 -std=c99 -DNDEBUG -fPIC -shared -I"$REF/c" -I"$REF/c/subprojects/kastore" \
        $SRCS "$HERE/ref_shim.c" -lm -o "$OUT"
    echo "built $OUT"
fi
