#!/bin/bash
# TEST INFRASTRUCTURE, CPU only: data-race detection over the emulated device code.  The product sources are compiled with
# ThreadSanitizer and every CUDA thread of a CTA is a TSan fiber that is synchronised only by what the hardware
# synchronises (tests/emul/cuda_emul.h, "RACE DETECTION"); tests/emul/race_check.cpp drives every device path.
# The closest thing to `compute-sanitizer --tool racecheck` without a GPU.  Exit status 66 = at least one race reported.
#   bash tools/emul_race.sh [case ...]
cd "$(dirname "$0")/.."
B=$PWD/tests/emul/_build
mkdir -p $B
FLAGS="-x c++ -std=c++17 -O1 -g -fsanitize=thread --param tsan-distinguish-volatile=1 -pthread -DKR_EMUL -DKR_EMUL_TSAN -DKR_ENC_PIPE_MIN=256 -Itests/emul -Wno-unknown-pragmas"
g++ $FLAGS -o $B/race_check tests/emul/race_check.cpp || exit 1
# one identity per warp (see tsan_identity in cuda_emul.h): 32 fibers share a ThreadSanitizer thread, so no shadow call stack
g++ $FLAGS --param tsan-instrument-func-entry-exit=0 -o $B/race_check_warp tests/emul/race_check.cpp || exit 1
# the score table of the index scan shrunk to 64 slots: nearly every query overflows into 4, 16, ... passes (redo / dry rounds)
g++ $FLAGS -DKR_IXQ_H=64 -o $B/race_check_h64 tests/emul/race_check.cpp || exit 1
export KR_EMUL_SMS=${KR_EMUL_SMS:-2}
export TSAN_OPTIONS="halt_on_error=0 exitcode=66 history_size=4 report_signal_unsafe=0 ${TSAN_EXTRA}"
if [ $# -gt 0 ]; then $B/race_check "$@"; exit $?; fi
W="env KR_EMUL_RACE_ID=warp $B/race_check_warp"
# positive controls first: a detector that cannot see a seeded bug proves nothing by staying silent
for c in ctl_syncthreads ctl_syncwarp ctl_global ctl_host_early ctl_two_streams ctl_free_order; do
    $B/race_check $c > $B/race_$c.log 2>&1
    [ $? -eq 66 ] || { echo "emul_race: control $c was NOT reported"; exit 2; }
done
for c in ctl_syncthreads ctl_host_early ctl_two_streams ctl_free_order; do      # (conflicts between lanes of one warp are invisible per warp, by construction)
    $W $c > $B/race_warp_$c.log 2>&1
    [ $? -eq 66 ] || { echo "emul_race: control $c was NOT reported with warp identities"; exit 2; }
done
NT=512 NG=4 $W ctl_two_streams_wide > $B/race_warp_wide.log 2>&1
[ $? -eq 66 ] || { echo "emul_race: the wide two-stream control (2 x 4 CTAs x 512 threads) was NOT reported with warp identities"; exit 2; }
for c in ctl_clean ctl_streams_clean; do
    $B/race_check $c > $B/race_$c.log 2>&1 || { echo "emul_race: the ordered control $c was reported"; head -40 $B/race_$c.log; exit 2; }
    $W $c > $B/race_warp_$c.log 2>&1 || { echo "emul_race: the ordered control $c was reported (warp identities)"; head -40 $B/race_warp_$c.log; exit 2; }
done
KR_EMUL_RACE=grid $W ctl_grid > $B/race_grid_ctl.log 2>&1
[ $? -eq 66 ] || { echo "emul_race: grid control was NOT reported"; exit 2; }
$W ctl_grid > $B/race_grid_ctl0.log 2>&1 || { echo "emul_race: grid control reported outside grid mode"; exit 2; }
KR_EMUL_RACE=grid $W ctl_clean > $B/race_grid_ctl1.log 2>&1 || { echo "emul_race: the clean control was reported in grid mode"; head -40 $B/race_grid_ctl1.log; exit 2; }
echo "emul_race: controls OK (seeded bugs reported at both granularities, ordered twins silent)"
rc=0
summary() { grep -A4 "WARNING: ThreadSanitizer" $1 | grep "#0" | sort | uniq -c | sort -rn | head -12 | cut -c1-200; }
CASES="matrix_k8 matrix_k10 matrix_k3 db_k8 shards multi bench_like misc"
for c in $CASES; do
    $B/race_check $c > $B/race_$c.log 2>&1; r=$?
    echo "emul_race: $c (thread identities: inside a CTA) exit $r, $(grep -c 'WARNING: ThreadSanitizer' $B/race_$c.log) report(s)"
    [ $r -eq 0 ] || { rc=$r; summary $B/race_$c.log; }
    $W $c > $B/race_warp_$c.log 2>&1; r=$?
    echo "emul_race: $c (warp identities: between warps, launches, streams, host) exit $r, $(grep -c 'WARNING: ThreadSanitizer' $B/race_warp_$c.log) report(s)"
    [ $r -eq 0 ] || { rc=$r; summary $B/race_warp_$c.log; }
    KR_EMUL_RACE=grid KR_EMUL_SMS=3 $W $c > $B/race_grid_$c.log 2>&1; r=$?
    echo "emul_race: $c (warp identities, CTAs of a launch concurrent) exit $r, $(grep -c 'WARNING: ThreadSanitizer' $B/race_grid_$c.log) report(s)"
    [ $r -eq 0 ] || { rc=$r; summary $B/race_grid_$c.log; }
done
for c in matrix_k3 db_k8; do
    $B/race_check_h64 $c > $B/race_h64_$c.log 2>&1; r=$?
    echo "emul_race: $c (64-slot score table, thread identities) exit $r, $(grep -c 'WARNING: ThreadSanitizer' $B/race_h64_$c.log) report(s)"
    [ $r -eq 0 ] || { rc=$r; summary $B/race_h64_$c.log; }
done
exit $rc
