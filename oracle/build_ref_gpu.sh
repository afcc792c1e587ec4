#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference PyTorch extension (module name
# "XbitOps", sources: src/dq_torch_ops.cc, src/cu/unpack_weight_2_to_7.cu,
# src/cu/gemv_w4a16_pt.cu -- reference setup.py:94-97) from a scratch copy of /root/reference
# for compute_100, and drops only the built .so under oracle/_ref/refgpu/ (git-ignored, but it
# travels to the GPU box).  No reference source is copied into this repository.
# The reference setup.py crashes on a GPU-less host unless CUDA_ARCH=ALL (setup.py:62-63).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref/refgpu"
[ -d "$REF" ] || { echo "no $REF: nothing to build"; exit 0; }
SCRATCH=$(mktemp -d /tmp/xbitops_ref_build.XXXXXX)
cp -r "$REF"/. "$SCRATCH"/
cd "$SCRATCH"
TORCH_CUDA_ARCH_LIST=10.0 CUDA_ARCH=ALL MAX_JOBS=4 python setup.py build_ext --inplace > "$SCRATCH/build.log" 2>&1 || { tail -50 "$SCRATCH/build.log"; exit 1; }
mkdir -p "$OUT"
cp XbitOps*.so "$OUT"/
echo "built: $(ls "$OUT")"
rm -rf "$SCRATCH"
