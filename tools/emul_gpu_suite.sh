#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: the `-m gpu` test files run against the emulator build of the library (tests/emul/), minus
# the tests that need CUDA tensors or B200-sized inputs.  What the first GPU call executes is then code whose logic AND
# whose test expectations have both been exercised.  ~20 min.
#   bash tools/emul_gpu_suite.sh [extra pytest args]
cd "$(dirname "$0")/.."
LIB=$PWD/tests/emul/_build/libkmersreps_emul.so
[ -f "$LIB" ] || g++ -x c++ -std=c++17 -O2 -fPIC -shared -pthread -DKR_EMUL -Itests/emul -Wno-unknown-pragmas -o "$LIB" kmers/reps_b200/csrc/kr_api.cu
KR_LIB_PATH=$LIB KR_EMUL_SMS=4 python -m pytest tests/test_gpu_parity.py tests/test_zz_subchunks.py -m gpu -q -p no:cacheprovider \
    -k "not at_scale and not full_size and not large_database and not buildsets" "$@"
