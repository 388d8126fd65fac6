#!/usr/bin/env bash
# Build libfastmat_b200.so (sm_100a) in-tree, incrementally (make + nvcc -MD dependency files under build/).
# `EMUL=1 ./build.sh` builds the host-emulation test library instead (tests/emul/libfmb_emul.so: same sources,
# -DFMB_EMULATE, kernel bodies run on host threads; never shipped).  `EXTRA_DEFS=-D... OUT=build/alt/x.so ./build.sh`
# builds an experiment variant into its own object directory.
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
  exit 0
fi
OUT=${OUT:-fastmat_b200/lib/libfastmat_b200.so}
OBJDIR=build/obj$(echo "${EXTRA_DEFS:-}" | tr -c 'A-Za-z0-9_\n' '_')
mkdir -p "$(dirname "$OUT")" "$OBJDIR"
UNITS="capi fft_engine fft_k_f32_pow2 fft_k_f32_gen fft_k_f64_pow2 fft_k_f64_gen fft_fast_f32_L6 fft_fast_f64_L6 fft_fast_f32_L7 fft_fast_f64_L7 fft_fast_f32_L8 fft_fast_f32_L9 fft_fast_f32_L10 fft_fast_f32_L11 fft_fast_f32_L12 fft_v32_a fft_v32_b fft_v32_m fft_v32_c fft_v32_1 fft_v32p_f fft_v32p_c0 fft_v32p_c1 fft_v32t fft_fast_f64_L8 fft_fast_f64_L9 fft_fast_f64_L10 fft_fast_f64_L11 fwht elementwise"
OBJS=""
for f in $UNITS; do OBJS="$OBJS $OBJDIR/$f.o"; done
cat > "$OBJDIR/Makefile" <<EOF
NVCC := $NVCC
FLAGS := $COMMON ${PTXAS_V:+-Xptxas -v}
$OUT: $OBJS $OBJDIR/planner.o
	\$(NVCC) -shared -o \$@ \$^ -lcudart_static -lpthread -ldl -lrt
$OBJDIR/%.o: $SRC/%.cu
	\$(NVCC) \$(FLAGS) -gencode arch=compute_100a,code=sm_100a -MD -MF $OBJDIR/\$*.d -c \$< -o \$@
$OBJDIR/planner.o: $SRC/planner.cpp
	\$(NVCC) \$(FLAGS) -gencode arch=compute_100a,code=sm_100a -MD -MF $OBJDIR/planner.d -c \$< -o \$@
-include $OBJDIR/*.d
EOF
make -s -j"${JOBS:-$(nproc)}" -f "$OBJDIR/Makefile" "$OUT"
echo "built $OUT"
