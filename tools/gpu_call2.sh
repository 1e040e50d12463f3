#!/bin/bash
# Second GPU call of a round (after tools/gpu_call1.sh passed): the ncu evidence of the committed revision.
#   gpurun --timeout 1500 -- 'bash tools/gpu_call2.sh'
# ncu is the only tool of this call; every profiled command line has exited 0 without ncu directly before (&&, no pipe).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
echo "=== launch list (cold-cache, serialised: shares, not absolutes)"
$CMD > gpurun_out/c2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $CMD > gpurun_out/c2_ncu_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c2_ncu_launches.log
echo "=== --set full: the scan (dominant kernel), then one launch each of the encoder and the index build's sort / scatter"
$CMD > gpurun_out/c2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ix_query -s 1 -c 1 -o gpurun_out/r02_prof_k_ix_query $CMD > gpurun_out/c2_ncu_full1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c2_ncu_full1.log
$CMD > gpurun_out/c2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_encode_sort_warp|k_ix_sort|k_ix_partition' -s 4 -c 4 -o gpurun_out/r02_prof_k1_k4 $CMD > gpurun_out/c2_ncu_full2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c2_ncu_full2.log
JOIN="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --path join"
$JOIN > gpurun_out/c2_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_join -s 1 -c 1 -o gpurun_out/r02_prof_k_join $JOIN > gpurun_out/c2_ncu_full3.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c2_ncu_full3.log
for f in r02_prof_k_ix_query r02_prof_k1_k4 r02_prof_k_join; do
    [ -f gpurun_out/$f.ncu-rep ] && ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
