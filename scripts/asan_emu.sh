#!/bin/sh
# AddressSanitizer pass over the CPU-emulated kernels (compute-sanitizer is closed on the GPU pool): builds the emulator
# library with -fsanitize=address under /tmp and runs the batched-affine MSM rounds (windowed, fixed-base, one heavily
# loaded bucket) and multi-pass NTTs against the oracle.  usage: sh scripts/asan_emu.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/zkp_asan
mkdir -p "$OUT"
for f in api ntt msm msm_affine gen poly sort; do
  g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -DZKP_EMU -fPIC -pthread -I"$ROOT/tests/emu" \
      -I"$ROOT/zkp-implementation_b200/csrc" -x c++ -c "$ROOT/zkp-implementation_b200/csrc/$f.cu" -o "$OUT/$f.o" &
done
wait
for f in plonk kzg transcript_api; do
  g++ -O1 -g -fsanitize=address -std=c++17 -fPIC -fopenmp -c "$ROOT/zkp-implementation_b200/host/$f.cpp" -o "$OUT/$f.host.o" &
done
wait
g++ -shared -pthread -fopenmp -fsanitize=address -o "$OUT/libzkp_b200_emu.so" "$OUT"/*.o
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python "$ROOT/scripts/asan_emu_target.py" "$OUT/libzkp_b200_emu.so"
