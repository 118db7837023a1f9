#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  The reference's search kernels (DV-Kernel.cu, unmodified, from $REF) compiled for
# sm_100a into oracle/_ref/libref_search_cuda.so: bench.py's "kernel to beat" on the same GPU.  Kept apart from
# build_ref.sh because the deep __forceinline__ recursion of DV-Kernel.cu takes cicc/ptxas very long;
# run it once in the background.  S3_REF_CUDA_OPT (default -O3) is passed to cicc and ptxas.  Measured on B200: the -O1
# build (one minute) runs the bench sample at 9.7 M reads/s, the -O3 build (40 minutes) at 3.7 M reads/s, so
# __graft_entry__.build() makes the -O1 one.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF" ] || { echo "[build_ref_search_cuda] $REF not present" >&2; exit 0; }
mkdir -p "$OUT"
NVCC=${S3_NVCC:-/usr/local/cuda/bin/nvcc}
OPT=${S3_REF_CUDA_OPT:--O3}
TARGET=${1:-$OUT/libref_search_cuda.so}
time $NVCC -w -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -Xcicc $OPT -Xptxas $OPT \
    -I"$REF" "$HERE/ref_shim/ref_search_cuda.cu" -o "$TARGET.tmp" -lcudart
mv "$TARGET.tmp" "$TARGET"
echo "nvcc -Xcicc $OPT -Xptxas $OPT" > "$TARGET.flags"
echo "[build_ref_search_cuda] $TARGET OK"
