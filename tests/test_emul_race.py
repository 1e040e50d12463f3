"""TEST INFRASTRUCTURE, CPU only: data-race detection over the emulated device code (tests/emul/race_check.cpp; the long
form is tools/emul_race.sh).  The product sources are compiled with ThreadSanitizer; every CUDA thread of a CTA is a TSan
fiber ordered only by what the hardware orders (barriers, __syncwarp, mbarriers), and every stream has a clock of its own
(host and streams are ordered only by synchronisations and events), so a missing barrier in a kernel, a host read before
the synchronisation, or two streams on one buffer without an event is a report -- the stand-in for `compute-sanitizer --tool racecheck` in a container without a GPU.  Here: the detector's own
controls (three seeded races must be reported, the ordered twin must not) and a quick pass over both device paths of the
matrix in both race modes; the full set of cases takes ~15 minutes and lives in the script."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "emul", "_build")
BIN = os.path.join(BUILD, "race_check_suite")
TSAN = "halt_on_error=0 exitcode=66 history_size=4 report_signal_unsafe=0"


FLAGS = ["-x", "c++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "--param", "tsan-distinguish-volatile=1", "-pthread",
         "-DKR_EMUL", "-DKR_EMUL_TSAN", "-DKR_ENC_PIPE_MIN=256", "-I" + os.path.join(ROOT, "tests", "emul"), "-Wno-unknown-pragmas"]


def build(path, extra):
    srcs = [os.path.join(ROOT, "tests", "emul", f) for f in ("race_check.cpp", "cuda_emul.h")]
    csrc = os.path.join(ROOT, "kmers", "reps_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        p = subprocess.run(["g++"] + FLAGS + extra + ["-o", path, os.path.join(ROOT, "tests", "emul", "race_check.cpp")],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert p.returncode == 0, p.stdout.decode()[-3000:]


@pytest.fixture(scope="module")
def race_bin():
    """two builds: one ThreadSanitizer identity per CUDA thread (conflicts inside a CTA, a missing __syncwarp included), and
    one per WARP (32 fibers share an identity, hence no shadow call stack): a few dozen identities instead of hundreds,
    which is what keeps conflicts between launches, streams and CTAs visible -- the detector has 256 thread slots"""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    os.makedirs(BUILD, exist_ok=True)
    build(BIN, [])
    build(BIN + "_warp", ["--param", "tsan-instrument-func-entry-exit=0"])
    # ThreadSanitizer needs an address-space layout it knows; where it cannot start at all the test has nothing to say
    probe = subprocess.run([BIN, "none"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(os.environ, TSAN_OPTIONS=TSAN))
    if probe.returncode != 0 and b"unexpected memory mapping" in probe.stdout:
        pytest.skip("ThreadSanitizer cannot run in this container: " + probe.stdout.decode()[-200:])
    return BIN


def run(binary, case, grid=False, warp=False, **extra):
    env = dict(os.environ, TSAN_OPTIONS=TSAN, KR_EMUL_SMS="2", **extra)
    if grid:
        env["KR_EMUL_RACE"] = "grid"
    if warp:
        env["KR_EMUL_RACE_ID"] = "warp"
        binary += "_warp"
    p = subprocess.run([binary, case], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=900)
    return p.returncode, p.stdout.decode()


def test_detector_controls(race_bin):
    # missing __syncthreads / missing __syncwarp (a shuffle orders nothing) / two threads on one global word /
    # host read of a pinned buffer before the synchronisation / two streams on one buffer without an event /
    # a stream-ordered free not ordered after the last use
    for ctl in ("ctl_syncthreads", "ctl_syncwarp", "ctl_global", "ctl_host_early", "ctl_two_streams", "ctl_free_order"):
        rc, out = run(race_bin, ctl)
        assert rc == 66 and "WARNING: ThreadSanitizer: data race" in out, "seeded bug %s was not reported\n%s" % (ctl, out[-1500:])
        if ctl not in ("ctl_syncwarp", "ctl_global"):       # (conflicts between lanes of one warp are invisible per warp, by construction)
            rc, out = run(race_bin, ctl, warp=True)
            assert rc == 66, "seeded bug %s was not reported with warp identities\n%s" % (ctl, out[-1500:])
    # two streams x 4 CTAs x 512 threads on one buffer: far more CUDA threads than the detector has slots
    rc, out = run(race_bin, "ctl_two_streams_wide", warp=True, NT="512", NG="4")
    assert rc == 66, "the wide two-stream conflict was not reported with warp identities\n" + out[-1500:]
    for ctl in ("ctl_clean", "ctl_streams_clean"):
        for warp in (False, True):
            rc, out = run(race_bin, ctl, warp=warp)
            assert rc == 0, "the ordered control %s was reported:\n%s" % (ctl, out[-3000:])
    rc, out = run(race_bin, "ctl_grid", grid=True, warp=True)
    assert rc == 66, "a conflict between two CTAs was not reported in grid mode\n" + out[-1500:]
    rc, out = run(race_bin, "ctl_grid", warp=True)
    assert rc == 0, out[-3000:]
    rc, out = run(race_bin, "ctl_clean", grid=True, warp=True)
    assert rc == 0, out[-3000:]


@pytest.mark.parametrize("mode", ["thread", "warp", "warp-grid"])
def test_device_code_has_no_data_race_quick(race_bin, mode):
    """encode (counting sort, register network, global-memory sort; pipelined over the copy stream), block lists + TMA-fed
    tile join, index build beside the zero-fill on its own stream + the three scans, pairs: both device paths on one small
    batch, results cross-checked, no report"""
    rc, out = run(race_bin, "smoke", grid=mode.endswith("grid"), warp=mode.startswith("warp"))
    assert rc == 0, out[-4000:]
    assert "race_check: done" in out
