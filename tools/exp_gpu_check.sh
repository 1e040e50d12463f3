#!/bin/bash
# First GPU call after the round in which GPU access was closed: validate and time everything that was written
# against the CPU emulator only.  Run from the repo root on the GPU box, e.g.
#   gpurun --timeout 1500 -- 'bash tools/exp_gpu_check.sh'
# Results land in gpurun_out/exp_*.  Nothing here changes the default library.
set -x
mkdir -p gpurun_out
EXP=$PWD/kmers/reps_b200/libkmersreps_cuda_exp.so
# 0. every experimental path once, small, under compute-sanitizer: an out-of-bounds access or a shared-memory race is
#    reported here instead of faulting the device later (stop if memcheck finds anything)
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
KR_LIB_PATH=$EXP timeout 900 $CS --tool memcheck --error-exitcode 86 python tools/exp_sanitize.py > gpurun_out/exp_memcheck.log 2>&1
rc=$?; tail -5 gpurun_out/exp_memcheck.log
if [ $rc -ne 0 ]; then echo "memcheck failed (rc=$rc): not running the experimental paths at scale"; exit 1; fi
KR_LIB_PATH=$EXP timeout 900 $CS --tool racecheck python tools/exp_sanitize.py > gpurun_out/exp_racecheck.log 2>&1
tail -5 gpurun_out/exp_racecheck.log
# 1. the default library: the whole GPU suite (includes tests never run on hardware: neighbors, buildsets, sub-chunks)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/exp_pytest_default.log
# 2. the experimental library: the same suite + tests/test_zzz_experimental.py
EXPDIR=$(mktemp -d); ln -s $EXP $EXPDIR/libkmersreps_cuda.so     # the host tools find the library through LD_LIBRARY_PATH
KR_LIB_PATH=$EXPDIR/libkmersreps_cuda.so LD_LIBRARY_PATH=$EXPDIR timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/exp_pytest_exp.log
# 2b. the library with the experimental paths as DEFAULTS for every context (what the default becomes after this check):
#     the whole suite again, host tools included (they find the library through LD_LIBRARY_PATH)
V2DIR=$(mktemp -d); ln -s $PWD/kmers/reps_b200/libkmersreps_cuda_v2.so $V2DIR/libkmersreps_cuda.so
KR_LIB_PATH=$V2DIR/libkmersreps_cuda.so LD_LIBRARY_PATH=$V2DIR timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/exp_pytest_v2.log
KR_LIB_PATH=$V2DIR/libkmersreps_cuda.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --sparse-e2e > gpurun_out/exp_bench_v2.json 2> gpurun_out/exp_bench_v2.err
# 3. timings: default, then each experimental path
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/exp_bench_default.json 2> gpurun_out/exp_bench_default.err
KR_LIB_PATH=$EXP timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --exp-encode > gpurun_out/exp_bench_encode.json 2> gpurun_out/exp_bench_encode.err
KR_LIB_PATH=$EXP timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --exp-encode --exp-pairs --sparse-e2e > gpurun_out/exp_bench_pairs.json 2> gpurun_out/exp_bench_pairs.err
KR_LIB_PATH=$EXP timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --exp-encode --exp-pairs --exp-buckets --sparse-e2e > gpurun_out/exp_bench_buckets.json 2> gpurun_out/exp_bench_buckets.err
KR_LIB_PATH=$EXP timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --sparse-e2e > gpurun_out/exp_bench_sparse_join.json 2> gpurun_out/exp_bench_sparse_join.err
timeout 600 python bench.py --workload findclosest --steps 2 > gpurun_out/exp_fc_default.json 2> gpurun_out/exp_fc_default.err
KR_LIB_PATH=$EXP timeout 600 python bench.py --workload findclosest --steps 2 --exp-encode --exp-idxq --exp-gidx > gpurun_out/exp_fc_exp.json 2> gpurun_out/exp_fc_exp.err
tail -c 600 gpurun_out/exp_pytest_default.log gpurun_out/exp_pytest_exp.log gpurun_out/exp_pytest_v2.log
for f in gpurun_out/exp_bench_*.json gpurun_out/exp_fc_*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "stages_ms_per_step")}, (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e_sparse") or {}).get("ms_per_step"))
except Exception as e:
    print("no result:", e)
PY
done
