#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: the `-m gpu` test files run against the emulator build of the library (tests/emul/), with
# the B200-sized inputs of three tests shrunk.  Device tensors of the tests become host memory, the host commands are linked
# against the emulator library.  What the first GPU call executes is then code whose logic AND whose test expectations
# have both been exercised.  ~25 min.
#   bash tools/emul_gpu_suite.sh [extra pytest args]
cd "$(dirname "$0")/.."
B=$PWD/tests/emul/_build
mkdir -p $B
LIB=$B/libkmersreps_emul_suite.so
g++ -x c++ -std=c++17 -O2 -fPIC -shared -pthread -DKR_EMUL -Itests/emul -Wno-unknown-pragmas -o "$LIB" kmers/reps_b200/csrc/kr_api.cu || exit 1
g++ -O2 -std=c++17 -o $B/kmers_reps_emul kmers/reps_b200/host/tools.cpp "$LIB" -Wl,-rpath,$B -pthread || exit 1
# the three B200-sized tests run 40 x smaller (KR_TEST_SCALE): their own index arithmetic and invariants get exercised
KR_TEST_SCALE=40 KR_TOOL_PATH=$B/kmers_reps_emul KR_LIB_PATH=$LIB KR_EMUL_SMS=4 python -m pytest tests/test_gpu_parity.py tests/test_host_tools.py \
    tests/test_zz_subchunks.py -m gpu -q -p no:cacheprovider "$@"
