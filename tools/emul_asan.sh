#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: the product's device + host sources on the emulator under AddressSanitizer (the closest
# thing to compute-sanitizer memcheck a container without a GPU offers: every out-of-range global-memory access of a
# kernel is a heap overflow of a posix_memalign'd "device" allocation).
#   bash tools/emul_asan.sh [case ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/emul/_build
LIB=tests/emul/_build/libkmersreps_emul_asan.so
g++ -x c++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -shared -pthread -DKR_EMUL -Itests/emul \
    -Wno-unknown-pragmas -o $LIB kmers/reps_b200/csrc/kr_api.cu
CASES=${@:-matrix_layouts k10_rectangular_pairs find_closest_and_build subchunks sharded_import api_misc sparse_matrix index_crowded encode_warp}
for c in $CASES; do
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        KR_LIB_PATH=$PWD/$LIB KR_EMUL_SMS=2 python tests/emul/run_emul_checks.py $c 2>&1 | tail -3
done
