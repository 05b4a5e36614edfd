#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's vendored BLIS (third_party/blis, pinned at e8566eb) from a
# writable scratch copy and installs ONLY build outputs (libblis.a + the generated
# umbrella header blis.h) into oracle/_ref/blis/.  No reference source is copied into
# the repository: the scratch copy lives in $TMPDIR and is deleted afterwards.
#
# The reference tree is read-only and BLIS's configure writes into its source dir, hence
# the scratch copy.  Configuration = what the reference's own
# tests use: `./configure -t pthreads generic` (generic sub-config: KC=256 for float,
# third_party/blis/ref_kernels/bli_cntx_ref.c:382-384).
set -euo pipefail
REF=${LOF_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref/blis"
if [[ -f "$OUT/lib/libblis.a" && -f "$OUT/include/blis/blis.h" ]]; then
  echo "[oracle] BLIS already built at $OUT"; exit 0
fi
if [[ ! -d "$REF/third_party/blis" ]]; then
  echo "[oracle] no reference tree at $REF; cannot build BLIS" >&2; exit 3
fi
SCRATCH="$(mktemp -d "${TMPDIR:-/tmp}/lof_blis.XXXXXX")"
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF/third_party/blis" "$SCRATCH/blis"
cd "$SCRATCH/blis"
chmod -R u+w .
./configure -t pthreads --disable-shared --prefix="$SCRATCH/install" generic > "$SCRATCH/configure.log" 2>&1
make -j"$(nproc)" > "$SCRATCH/make.log" 2>&1
make install > "$SCRATCH/install.log" 2>&1
mkdir -p "$OUT/lib" "$OUT/include"
cp "$SCRATCH/install/lib/libblis.a" "$OUT/lib/"
# generated umbrella headers (include/blis.h is a stub that includes blis/blis.h)
cp -r "$SCRATCH/install/include/." "$OUT/include/"
echo "[oracle] BLIS built -> $OUT"
