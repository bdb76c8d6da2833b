#!/bin/bash
# Build an A/B variant of libtmf.so into variants/ (scratch, git-ignored, travels to the GPU box with gpurun):
#   scripts/build_variant.sh NAME [FILE.cu] [extra nvcc flags, e.g. -DSOME_SWITCH=1]
# recompiles FILE.cu (default train.cu) with the extra flags, links it with the other objects of the last `make`, and writes
# variants/libtmf_NAME.so.  `VARIANTS="base NAME" bash scripts/gpu_trip_ab.sh` then benches each variant on ONE box
# (same GPU, same clocks) -- run-to-run differences between boxes are larger than most kernel tweaks.
set -e
NAME=$1; shift
SRC=train.cu
if [[ "$1" == *.cu ]]; then SRC=$1; shift; fi
cd "$(dirname "$0")/../teamoflow_b200/csrc"
make -s
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v"
mkdir -p ../../variants
$NV "$@" -c $SRC -o /tmp/variant_$NAME.o 2> ../../variants/$NAME.ptxas.log
OBJS=""
for f in setup train score score_topk peer gemm_tc; do
  if [ "$f.cu" == "$SRC" ]; then OBJS="$OBJS /tmp/variant_$NAME.o"; else OBJS="$OBJS $f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libtmf_$NAME.so $OBJS
echo "variants/libtmf_$NAME.so built (ptxas log: variants/$NAME.ptxas.log)"
