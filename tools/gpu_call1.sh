#!/bin/bash
# First GPU call of a round: prove the DEFAULT library cannot fault before anything large runs, then regain the records.
#   gpurun --timeout 2400 -- 'bash tools/gpu_call1.sh'
# Stops at the first failing stage.  Exactly one sanitizer tool (memcheck) is used in the call.
mkdir -p gpurun_out
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
stage() { echo "=== $1"; }
fail() { echo "STAGE FAILED: $1"; tail -40 "$2"; exit 1; }

stage "1. memcheck: smoke() with the default library"
timeout 900 $CS --tool memcheck --error-exitcode 86 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_memcheck_smoke.log 2>&1 || fail memcheck-smoke gpurun_out/c1_memcheck_smoke.log
tail -3 gpurun_out/c1_memcheck_smoke.log

stage "2. memcheck: N-shard import on one GPU + reference fixtures (small)"
timeout 1200 $CS --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
    -k "shard_import or kat_k10 or vapb_matrix or config1_find_closest or sets_from_key_lists or error_paths or (multi_handle and index)" > gpurun_out/c1_memcheck_tests.log 2>&1 || fail memcheck-tests gpurun_out/c1_memcheck_tests.log
tail -3 gpurun_out/c1_memcheck_tests.log

stage "3. the whole GPU suite, default library, no tool"
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c1_pytest_gpu.log 2>&1 || fail pytest-gpu gpurun_out/c1_pytest_gpu.log
tail -3 gpurun_out/c1_pytest_gpu.log

stage "4. bench.py (N=1, defaults) + the reference arm"
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err || fail bench gpurun_out/c1_bench.err
tail -c 1500 gpurun_out/c1_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err || fail bench-ref gpurun_out/c1_bench_ref.err
tail -c 600 gpurun_out/c1_bench_ref.json

stage "4b. the tile join pinned, for the comparison of the two device paths"
timeout 600 python bench.py --path join --steps 5 --no-cpu > gpurun_out/c1_bench_join.json 2> gpurun_out/c1_bench_join.err || fail bench-join gpurun_out/c1_bench_join.err
tail -c 700 gpurun_out/c1_bench_join.json

stage "5. findClosest / greedy build workload (configs 3, 4), default library"
timeout 1200 python bench.py --workload findclosest --steps 2 > gpurun_out/c1_fc.json 2> gpurun_out/c1_fc.err || fail bench-fc gpurun_out/c1_fc.err
tail -c 1500 gpurun_out/c1_fc.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "ALL STAGES OK"
