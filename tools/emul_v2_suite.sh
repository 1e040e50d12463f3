#!/bin/bash
# CPU-only: the GPU test files against an EMULATOR build in which every experimental path is the default
# (-DKR_DEFAULTS_V2).  Takes ~4 minutes; not part of the pytest CPU suite for that reason.  Run from the repo root.
set -e
D=$(mktemp -d)
g++ -x c++ -std=c++17 -O2 -fPIC -shared -pthread -DKR_EMUL -DKR_EXPERIMENTAL -DKR_DEFAULTS_V2 -Itests/emul -Wno-unknown-pragmas \
    -o $D/libkmersreps_cuda.so kmers/reps_b200/csrc/kr_api.cu
KR_V2_SELFCHECK=0 KR_LIB_PATH=$D/libkmersreps_cuda.so LD_LIBRARY_PATH=$D KR_EMUL_SMS=4 python -m pytest tests/test_gpu_parity.py tests/test_host_tools.py \
    tests/test_zz_subchunks.py -m gpu -q -p no:cacheprovider \
    --deselect tests/test_gpu_parity.py::test_index_path_multi_slab_large_batch \
    --deselect tests/test_gpu_parity.py::test_synthetic_families_vs_oracle \
    --deselect tests/test_gpu_parity.py::test_greedy_build_properties_at_scale
