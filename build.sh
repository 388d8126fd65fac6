#!/usr/bin/env bash
# Build libfastmat_b200.so (sm_100a) in-tree.  `EMUL=1 ./build.sh` builds the host-emulation test library instead
# (tests/emul/libfmb_emul.so: same sources, -DFMB_EMULATE, kernel bodies run on host threads; never shipped).
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")"
SRC=fastmat_b200/csrc
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
COMMON="${EXTRA_DEFS:-} -std=c++20 -O3 --extended-lambda -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas -Xcompiler -Wno-unused-but-set-variable -Xcompiler -Wno-unused-function -Iinclude --expt-relaxed-constexpr -diag-suppress 177,550"
if [ "${EMUL:-0}" = "1" ]; then
  mkdir -p tests/emul
  $NVCC $COMMON -DFMB_EMULATE -gencode arch=compute_100a,code=sm_100a -shared -o tests/emul/libfmb_emul.so \
     $SRC/capi.cu $SRC/fft_engine.cu $SRC/fwht.cu $SRC/elementwise.cu $SRC/planner.cpp $SRC/emulate.cpp -lpthread
  echo "built tests/emul/libfmb_emul.so"
else
  mkdir -p fastmat_b200/lib build
  OBJS=""
  for f in capi fft_engine fft_k_f32_pow2 fft_k_f32_gen fft_k_f64_pow2 fft_k_f64_gen fft_fast_f32_L8 fft_fast_f32_L9 fft_fast_f32_L10 fft_fast_f32_L11 fft_fast_f32_L12 fft_v32_a fft_v32_b fft_v32_m fft_v32_c fft_v32p_f fft_v32p_c0 fft_v32p_c1 fft_fast_f64_L8 fft_fast_f64_L9 fft_fast_f64_L10 fft_fast_f64_L11 fft_fused_f32_8_8 fft_fused_f32_8_9 fft_fused_f32_9_9 fft_fused_f32_9_10 fft_fused_f32_10_10 fft_fused_f32_10_11 fft_fused_f32_11_11 fft_fused_f64_8_8 fwht elementwise; do
    $NVCC $COMMON ${PTXAS_V:+-Xptxas -v} -gencode arch=compute_100a,code=sm_100a -c $SRC/$f.cu -o build/$f.o &
    OBJS="$OBJS build/$f.o"
  done
  $NVCC $COMMON -c $SRC/planner.cpp -o build/planner.o &
  wait
  $NVCC -shared -o fastmat_b200/lib/libfastmat_b200.so $OBJS build/planner.o -lcudart_static -lpthread -ldl -lrt
  echo "built fastmat_b200/lib/libfastmat_b200.so"
fi
