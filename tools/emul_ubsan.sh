#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: the product's device + host sources on the emulator under UndefinedBehaviorSanitizer
# (alignment + array bounds).  A typed access that is not aligned to its type -- a uint4 store at an address that is not a
# multiple of 16, a 64-bit entry at an odd word -- is a "misaligned address" device fault on hardware; x86 would let it pass.
#   bash tools/emul_ubsan.sh [case ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/emul/_build
LIB=tests/emul/_build/libkmersreps_emul_ub.so
g++ -x c++ -std=c++17 -O1 -g -fsanitize=alignment,bounds -fno-sanitize-recover=all -fPIC -shared -pthread -DKR_EMUL -Itests/emul \
    -Wno-unknown-pragmas -o $LIB kmers/reps_b200/csrc/kr_api.cu
CASES=${@:-matrix_layouts k10_rectangular_pairs find_closest_and_build subchunks sharded_import api_misc sparse_matrix index_crowded encode_warp set_ops multi}
for c in $CASES; do
    LD_PRELOAD=$(gcc -print-file-name=libubsan.so) KR_LIB_PATH=$PWD/$LIB KR_EMUL_SMS=2 python tests/emul/run_emul_checks.py $c 2>&1 | tail -3
done
