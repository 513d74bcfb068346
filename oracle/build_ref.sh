#!/usr/bin/env bash
# Build the UNMODIFIED reference (/root/reference/gpssim.c) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
#
# The reference sources are compiled where they lie; nothing is copied into the
# tracked tree.  Outputs (git-ignored, but shipped to the GPU box by gpurun):
#
#   oracle/_ref/gps-sdr-sim-float      reference exactly as shipped
#                                      (gpssim.h:4 defines FLOAT_CARR_PHASE)
#   oracle/_ref/gps-sdr-sim-int        same gpssim.c, gpssim.h:4 disabled - the
#                                      integer-carrier branch (gpssim.c:2202, :2176,
#                                      :2252, :1624-1625) the north-star targets
#   oracle/_ref/libgpssim_ref_{int,float}.so
#                                      the same translation unit as a shared object
#                                      (main renamed) so tests can call codegen(),
#                                      computeChecksum(), read sinTable512[] ... via ctypes
#   oracle/_ref/data/                  scenario input files named in BASELINE.json
#
# Flags are the reference Makefile's (Makefile:8,12): -O3 -Wall -D_FILE_OFFSET_BITS=64, -lm.
# No -march / -ffast-math: they change the output bytes.
set -euo pipefail

REF="${GPSSIM_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"

if [ ! -f "$REF/gpssim.c" ]; then
    echo "build_ref.sh: $REF/gpssim.c not found - keeping prebuilt oracle/_ref as is" >&2
    exit 0
fi

mkdir -p "$OUT/data"
# -ffp-contract=off: a no-op on x86-64 without -march (no FMA to contract into), spelled out so that the
# goldens mean the same thing on any build machine (see integration/build_host.py)
CFLAGS="-O3 -Wall -D_FILE_OFFSET_BITS=64 -ffp-contract=off"

# --- as shipped (double carrier phase) ---------------------------------------
gcc $CFLAGS "$REF/gpssim.c" -lm -o "$OUT/gps-sdr-sim-float"
gcc $CFLAGS -fPIC -shared -Dmain=gpssim_ref_main "$REF/gpssim.c" -lm -o "$OUT/libgpssim_ref_float.so"

# --- integer carrier phase: the one #define at gpssim.h:4 switched off --------
# The edited header only ever exists in a scratch directory.  gpssim.c is fed on
# stdin so that its '#include "gpssim.h"' resolves against the scratch copy.
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
sed 's|^#define FLOAT_CARR_PHASE|// &|' "$REF/gpssim.h" > "$TMP/gpssim.h"
if grep -q '^#define FLOAT_CARR_PHASE' "$TMP/gpssim.h"; then
    echo "build_ref.sh: failed to disable FLOAT_CARR_PHASE" >&2; exit 1
fi
( cd "$TMP" && gcc $CFLAGS -x c - -lm -o "$OUT/gps-sdr-sim-int" < "$REF/gpssim.c" )
( cd "$TMP" && gcc $CFLAGS -fPIC -shared -Dmain=gpssim_ref_main -x c - -lm \
      -o "$OUT/libgpssim_ref_int.so" < "$REF/gpssim.c" )

# --- scenario inputs -----------------------------------------------------------
for f in brdc3540.14n circle.csv satellite.csv rocket.csv triumphv3.txt; do
    cp -f "$REF/$f" "$OUT/data/$f"
    chmod u+w "$OUT/data/$f"
done

echo "build_ref.sh: built $(ls "$OUT" | tr '\n' ' ')"
