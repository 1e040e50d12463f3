#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: data-race detection over the emulated device code.  The product sources are compiled with
# ThreadSanitizer and every CUDA thread of a CTA is a TSan fiber that is synchronised only by what the hardware
# synchronises (tests/emul/cuda_emul.h, "RACE DETECTION"); tests/emul/race_check.cpp drives every device path.
# The closest thing to `compute-sanitizer --tool racecheck` without a GPU.  Exit status 66 = at least one race reported.
#   bash tools/emul_race.sh [case ...]
cd "$(dirname "$0")/.."
B=$PWD/tests/emul/_build
mkdir -p $B
g++ -x c++ -std=c++17 -O1 -g -fsanitize=thread --param tsan-distinguish-volatile=1 -pthread -DKR_EMUL -DKR_EMUL_TSAN -DKR_ENC_PIPE_MIN=256 -Itests/emul \
    -Wno-unknown-pragmas -o $B/race_check tests/emul/race_check.cpp || exit 1
export KR_EMUL_SMS=${KR_EMUL_SMS:-2}
export TSAN_OPTIONS="halt_on_error=0 exitcode=66 history_size=4 report_signal_unsafe=0 ${TSAN_EXTRA}"
if [ $# -gt 0 ]; then $B/race_check "$@"; exit $?; fi
# positive controls first: a detector that cannot see a missing barrier proves nothing by staying silent
for c in ctl_syncthreads ctl_syncwarp ctl_global ctl_host_early ctl_two_streams ctl_free_order; do
    $B/race_check $c > $B/race_$c.log 2>&1
    [ $? -eq 66 ] || { echo "emul_race: control $c was NOT reported"; exit 2; }
done
$B/race_check ctl_clean > $B/race_ctl_clean.log 2>&1 || { echo "emul_race: the clean control was reported"; cat $B/race_ctl_clean.log | head -40; exit 2; }
$B/race_check ctl_streams_clean > $B/race_ctl_streams_clean.log 2>&1 || { echo "emul_race: the ordered stream control was reported"; head -40 $B/race_ctl_streams_clean.log; exit 2; }
echo "emul_race: controls OK (6 seeded races reported, 2 ordered twins silent)"
rc=0
for c in matrix_k8 matrix_k10 matrix_k3 db_k8 shards multi bench_like; do
    $B/race_check $c > $B/race_$c.log 2>&1
    r=$?
    n=$(grep -c "WARNING: ThreadSanitizer" $B/race_$c.log)
    echo "emul_race: $c exit $r, $n report(s)"
    [ $r -eq 0 ] || { rc=$r; grep -A4 "WARNING: ThreadSanitizer" $B/race_$c.log | grep "#0\|#1" | sort | uniq -c | sort -rn | head -20; }
done
# grid mode: the CTAs of a launch are concurrent; a global word two CTAs touch without atomics is a race
KR_EMUL_RACE=grid $B/race_check ctl_grid > $B/race_grid_ctl.log 2>&1
[ $? -eq 66 ] || { echo "emul_race: grid control was NOT reported"; exit 2; }
$B/race_check ctl_grid > $B/race_grid_ctl0.log 2>&1 || { echo "emul_race: grid control reported outside grid mode"; exit 2; }
KR_EMUL_RACE=grid $B/race_check ctl_clean > $B/race_grid_ctl1.log 2>&1 || { echo "emul_race: the clean control was reported in grid mode"; head -40 $B/race_grid_ctl1.log; exit 2; }
echo "emul_race: grid controls OK"
for c in matrix_k8 matrix_k10 matrix_k3 db_k8 shards; do
    KR_EMUL_RACE=grid KR_EMUL_SMS=3 $B/race_check $c > $B/race_grid_$c.log 2>&1
    r=$?
    n=$(grep -c "WARNING: ThreadSanitizer" $B/race_grid_$c.log)
    echo "emul_race: $c (grid mode) exit $r, $n report(s)"
    [ $r -eq 0 ] || { rc=$r; grep -A4 "WARNING: ThreadSanitizer" $B/race_grid_$c.log | grep "#0\|#1" | sort | uniq -c | sort -rn | head -20; }
done
# the score table of the index scan shrunk to 64 slots: nearly every query overflows into 4, 16, ... passes (redo / dry rounds)
g++ -x c++ -std=c++17 -O1 -g -fsanitize=thread --param tsan-distinguish-volatile=1 -pthread -DKR_EMUL -DKR_EMUL_TSAN -DKR_IXQ_H=64 -Itests/emul \
    -Wno-unknown-pragmas -o $B/race_check_h64 tests/emul/race_check.cpp || exit 1
for c in matrix_k3 db_k8; do
    $B/race_check_h64 $c > $B/race_h64_$c.log 2>&1
    r=$?
    n=$(grep -c "WARNING: ThreadSanitizer" $B/race_h64_$c.log)
    echo "emul_race: $c (64-slot score table) exit $r, $n report(s)"
    [ $r -eq 0 ] || { rc=$r; grep -A4 "WARNING: ThreadSanitizer" $B/race_h64_$c.log | grep "#0\|#1" | sort | uniq -c | sort -rn | head -20; }
done
exit $rc
