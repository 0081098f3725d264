#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference CUDA rasterizer
# (/root/reference/hierslam-diff-gaussian-rasterization-w-depth) for sm_100 so
# that parity tests and bench.py can run it beside the new implementation on
# the GPU box.  Outputs go ONLY to oracle/_ref/S<S>/ (git-ignored, but shipped
# to the GPU box by gpurun).  No reference source is committed to this repo:
# the sources are copied to a scratch dir, NUM_SEMANTIC in the COPY's config.h
# is set the way the reference README (README.md:158-170) tells users to do,
# and the copy is compiled with torch's cpp_extension.
#
#   usage: oracle/build_ref.sh [S ...]        (default: 26)
#
# It also copies the reference's CALLERS of the rasterizer (scripts/hierslam.py and utils/*.py: plain torch code
# that needs a GPU at run time) to oracle/_ref/callers/ so that tests/test_reference_callers.py can replay
# get_loss_semantic / get_loss_semantic_mlp / add_new_gaussians_semantic_newrender UNCHANGED on the GPU box, once
# bound to the reference build and once to this repository's module.  /root/reference does not exist there.
#
# gcc-13 needs `-include cstdint` because rasterizer_impl.h uses uintptr_t /
# uint32_t without including <cstdint>; no source patch.
set -euo pipefail
REF=${HS_REFERENCE_ROOT:-/root/reference}/hierslam-diff-gaussian-rasterization-w-depth
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF" ]; then
  echo "reference not present at $REF - nothing to build (prebuilt oracle/_ref is used if it exists)"
  exit 0
fi
CALLERS=$OUT/callers
REFROOT=${HS_REFERENCE_ROOT:-/root/reference}
if [ ! -f "$CALLERS/scripts/hierslam.py" ]; then
  mkdir -p "$CALLERS/scripts" "$CALLERS/utils"
  cp "$REFROOT/scripts/hierslam.py" "$CALLERS/scripts/"
  cp "$REFROOT"/utils/*.py "$CALLERS/utils/"
  chmod -R u+w "$CALLERS"
  echo "copied the reference's rasterizer callers to oracle/_ref/callers"
fi
SVALS=("$@"); [ ${#SVALS[@]} -eq 0 ] && SVALS=(26)
for S in "${SVALS[@]}"; do
  DST=$OUT/S$S
  if ls "$DST"/diff_gaussian_rasterization/_C*.so >/dev/null 2>&1; then
    echo "oracle/_ref/S$S already built"; continue
  fi
  TMP=$(mktemp -d /tmp/hsref_S${S}_XXXX)
  cp -r "$REF"/. "$TMP"/
  sed -i -E "s/^#define NUM_SEMANTIC [0-9]+/#define NUM_SEMANTIC $S/" "$TMP/cuda_rasterizer/config.h"
  grep -q "#define NUM_SEMANTIC $S" "$TMP/cuda_rasterizer/config.h"
  ( cd "$TMP" && NVCC_APPEND_FLAGS="-include cstdint" TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS=${MAX_JOBS:-4} \
      python setup.py build_ext --inplace > "$TMP/build.log" 2>&1 ) || { tail -30 "$TMP/build.log"; exit 1; }
  mkdir -p "$DST/diff_gaussian_rasterization"
  cp "$TMP"/diff_gaussian_rasterization/_C*.so "$DST/diff_gaussian_rasterization/"
  cp "$TMP"/diff_gaussian_rasterization/__init__.py "$DST/diff_gaussian_rasterization/"
  echo "$S" > "$DST/NUM_SEMANTIC"
  rm -rf "$TMP"
  echo "built oracle/_ref/S$S"
done
