#!/usr/bin/env bash
# Builds the UNMODIFIED reference `gsplatcu` CUDA extension (gsplatcu/setup.py:4-14 of the
# reference tree) for sm_100a into baseline/_ref/ so that benchmarks/compare_ref_gpu.py can
# time it next to ours on the same B200 ("B-REF-GPU" in BASELINE.md).
# The reference tree is read-only, so the build runs from a scratch copy under /tmp; no
# reference source enters this repository (baseline/_ref/ is git-ignored, only the built .so
# lands there).  Only runs where /root/reference exists (the dev container).
set -euo pipefail
REF=${REF:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -d "$REF/gsplatcu" ] || { echo "no reference tree at $REF - skipping"; exit 0; }
TMP=$(mktemp -d /tmp/refgsplatcu.XXXXXX)
cp -r "$REF/gsplatcu" "$TMP/src"
mkdir -p "$OUT"
cd "$TMP/src"
TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS=4 python setup.py build_ext --build-lib "$OUT" --build-temp "$TMP/build" > "$TMP/build.log" 2>&1 \
  || { tail -30 "$TMP/build.log"; exit 1; }
ls -la "$OUT"
rm -rf "$TMP"
