"""bench.py end to end on the CPU emulator build of the library (tests/emul/): TEST INFRASTRUCTURE that runs the file the
driver runs -- argument handling, sharding, the oracle verification, the roofline / cpu_baseline / e2e bookkeeping and the
JSON contract -- at toy sizes in a container without a GPU.  The timings it prints mean nothing; the product library and
a real run still need a B200 (bench.py picks "cuda" + NCCL unless KR_LIB_PATH names the emulator build)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_emul_device_code import emul_lib  # noqa: E402,F401  (fixture: builds tests/emul/_build/libkmersreps_emul.so)

CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks")


def run(cmd, lib, timeout=1500, **env):
    e = dict(os.environ, KR_LIB_PATH=lib, KR_EMUL_SMS="2", **env)
    p = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    lines = [ln for ln in p.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout.decode()[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("path", ["index", "join"])
def test_allvsall_line(emul_lib, path):
    d = run([sys.executable, "bench.py", "--nseq", "600", "--steps", "1", "--warmup", "1", "--e2e-steps", "1", "--cpu-sample", "200", "--path", path], emul_lib)
    for k in CONTRACT:
        assert k in d, k
    assert d["config"]["workload"].startswith("all-vs-all") and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["traffic"] is None and r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] in (("k_ix_query<DENSE>", "k_zero_fill") if path == "index" else ("k_join<MATRIX>",))
    if path == "index":
        assert d["stages_ms_per_step"]["fill_ms"] > 0          # the batch is re-encoded every step: the fill runs beside the index build
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e_dense"]["d2h_bytes_per_step"] > d["e2e"]["d2h_bytes_per_step"] / 4
    if path == "index":
        assert d["index"]["postings_walked_by_a_self_scan"] > d["index"]["entries"] > 0


def test_allvsall_two_ranks(emul_lib):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    d = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
             "bench.py", "--gpus", "2", "--nseq", "700", "--steps", "1", "--warmup", "1", "--e2e-steps", "1", "--path", "index", "--gather-cells"],
            emul_lib, KR_EMUL_DEVICES="2")
    assert d["n_gpus"] == 2 and d["cpu_baseline"] is None and d["cells_allgather"]["cells_total"] > 0
    assert "2 rank(s)" in d["config"]["sharding"]


def test_findclosest_line(emul_lib):
    d = run([sys.executable, "bench.py", "--workload", "findclosest", "--ncand", "1500", "--nq", "800", "--nrep", "300", "--q-chunk", "500", "--steps", "1",
             "--warmup", "1", "--cpu-queries", "50", "--cpu-build", "600"], emul_lib)
    assert d["metric"].startswith("findClosest") and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["kernel"] == "k_ix_query<ARGMAX>" and r["M_q_mean"] > 0 and r["bytes_per_query"] == pytest.approx(r["n_q_mean"] * 24 + 4 * r["M_q_mean"] + 8)
    assert d["cpu_baseline"]["unit"] == "queries/s" and d["cpu_baseline_build"]["gpu_same_prefix"]["identical_result"] is True


def test_findclosest_two_ranks(emul_lib):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    d = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
             "bench.py", "--gpus", "2", "--workload", "findclosest", "--ncand", "1200", "--nq", "601", "--nrep", "250", "--q-chunk", "200", "--steps", "1",
             "--warmup", "1"], emul_lib, KR_EMUL_DEVICES="2")
    assert d["n_gpus"] == 2 and d["result_allgather_ms"] is not None and d["cpu_baseline"] is None and d["roofline"]["kernel"] == "k_ix_query<ARGMAX>"


def test_reference_arm():
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--nseq", "3000", "--ref-sample", "400"], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    d = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
