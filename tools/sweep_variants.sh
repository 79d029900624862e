#!/bin/bash
# Builds A/B variants of libsph_b200.so that differ in compile-time tunables of sph_gather.cu (block size, stack depth,
# list entries in flight) under build/variants/ (git-ignored, travels to the GPU box), for
#   SPH_B200_LIB=build/variants/<name>.so python tools/quick_bench.py C2_dambreak_1M 20 grid
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$HERE/fluid-simulation-3d_b200/csrc"
OUT="$HERE/build/variants"
mkdir -p "$OUT"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off --expt-relaxed-constexpr -ccbin g++"
build() {   # name, flags
    local name="$1"; shift
    $NV "$@" -c "$SRC/sph_gather.cu" -o "$OUT/$name.gather.o"
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin g++ -o "$OUT/$name.so" \
        "$SRC/sph_api.o" "$SRC/sph_kernels.o" "$OUT/$name.gather.o" "$SRC/sph_sort.o" "$SRC/sph_multi.o" -lcudart -ldl
    rm -f "$OUT/$name.gather.o"
    echo "built $name"
}
build t64 -DSPH_WALK_THREADS=64 &
build t256 -DSPH_WALK_THREADS=256 &
build pks16 -DSPH_PKS=16 &
build pks32 -DSPH_PKS=32 &
wait
build lu2 -DSPH_LIST_UNROLL=2 &
build lu6 -DSPH_LIST_UNROLL=6 &
build vu2 -DSPH_VISC_UNROLL=2 &
build vu6 -DSPH_VISC_UNROLL=6 &
wait
