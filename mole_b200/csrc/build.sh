#!/usr/bin/env bash
# Builds mole_b200/libmole_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libmole_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX=/usr/bin/g++
[ -x "$HOSTCXX" ] || HOSTCXX=g++
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin "$HOSTCXX"
       -Xptxas -v --expt-relaxed-constexpr)
mkdir -p "$HERE/_obj"
"$NVCC" "${FLAGS[@]}" -c "$HERE/mole_api.cu" -o "$HERE/_obj/mole_api.o" 2> "$HERE/_obj/ptxas_mole_api.log" || { cat "$HERE/_obj/ptxas_mole_api.log"; exit 1; }
"$NVCC" "${FLAGS[@]}" -x cu -c "$HERE/mole_host.cpp" -o "$HERE/_obj/mole_host.o"
"$NVCC" "${FLAGS[@]}" -x cu -c "$HERE/mole_comm.cpp" -o "$HERE/_obj/mole_comm.o"
"$NVCC" -shared -ccbin "$HOSTCXX" -o "$OUT" "$HERE/_obj/mole_api.o" "$HERE/_obj/mole_host.o" "$HERE/_obj/mole_comm.o" -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
