#!/usr/bin/env python
"""bench.py -- headline benchmark of the protein k-mer similarity hot path.

Workload (BASELINE.json configs[2], the one the metric is quoted on): all-vs-all 8-mer similarity of
100,000 synthetic ~300-aa seed proteins in 2,000 families (seed 0x5EED0003).  It fits one
GPU, so N=1 runs it whole; N>1 shards the 256-row stripes of the result over the ranks (strong scaling:
total work fixed; every rank encodes and indexes the whole batch and scans its own rows -- no data-path
collective).  A step = one pass of the hot path over the batch:
    residues (HBM) -> K1 encode + sort/dedupe -> K4 sorted inverted index -> K5 index scan that writes every
    cell of this rank's stripe of the uint16 similarity matrix (HBM)
(the library picks the index scan for this family-structured batch from the index's posting-list statistics;
--path join pins the tile join, the path of dense collections).
`value` = pairs (i<j) per second over all ranks, device time (CUDA events on the library's stream), max over
ranks.  `e2e` = the same through the host-buffer C ABI a caller uses (kr_encode + kr_sim_matrix_sparse64): H2D of
the residues and D2H of this rank's result inside the timed region, the result being the cells with sim >= 1 (all
other cells are 0 by definition); `e2e_dense` moves the dense stripe instead (kr_sim_matrix).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nseq 100000] [--path auto|join|index]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ensure_built():
    """build the native artefacts if this checkout has none (rank 0 builds, the others wait for the files)"""
    need = [os.path.join(ROOT, "kmers", "reps_b200", f) for f in ("libkmersreps_cuda.so", "libkrsynth.so")]
    if all(os.path.exists(p) for p in need):
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__ as g
        g.build()
    else:
        for _ in range(600):
            if all(os.path.exists(p) for p in need):
                break
            time.sleep(1)
        time.sleep(2)

METRIC = "protein k-mer pair comparisons/sec (all-vs-all)"
UNIT = "pairs/s"
SEED = 0x5EED0003


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload(nseq):
    from kmers.reps_b200 import synth
    nfam = max(1, nseq // 50)
    res, off, fam = synth.generate(SEED, nseq, nfam)
    return res, off, nfam


def cpu_baseline(res, off, sample, threads):
    """The oracle (a port of the reference's HashSet<String> algorithm) on the first `sample` sequences."""
    from oracle import kr_oracle as ko
    o = ko.get()
    n = min(sample, len(off) - 1)
    b = o.batch(res=res[: int(off[n])], off=off[: n + 1], K=8)
    t0 = time.perf_counter()
    pairs, _ = o.all_pairs_timed(b, ident_inf=True, threads=threads)
    dt = time.perf_counter() - t0
    return pairs / dt, pairs, dt, n


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm for this path.  The Java code cannot run here (no
    JVM; classes in un-vendored org.theseed:shared), so this times the oracle port on all host threads."""
    if rank != 0:
        return
    res, off, nfam = workload(args.nseq)
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    for _ in range(args.warmup):
        cpu_baseline(res, off, min(sample, 800), cores)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, pairs, dt, n = cpu_baseline(res, off, sample, cores)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "all-vs-all 8-mer similarity, %d synthetic ~300-aa seed proteins, %d families (BASELINE configs[2])" % (args.nseq, nfam),
                   "K": 8, "nseq": args.nseq, "cache": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "all i<j pairs of the first %d sequences of the workload (%d pairs) per step" % (n, pairs)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Dev:
    """Where device buffers live and how ranks talk: cuda:<local> + NCCL.  When KR_LIB_PATH points at the CPU emulator
    build of the library (tests/test_bench_emul.py: TEST INFRASTRUCTURE that exercises this file's logic in a container
    without a GPU; its timings mean nothing) "device" memory is host memory and the ranks use gloo."""

    def __init__(self, local):
        import torch
        import kmers.reps_b200 as kr
        self.torch = torch
        self.emul = "libkmersreps_emul" in os.path.basename(kr.lib_path())
        self.name = "cpu" if self.emul else "cuda"
        self.backend = "gloo" if self.emul else "nccl"
        self.local = local
        if not self.emul:
            torch.cuda.set_device(local)

    def init_dist(self, world):
        if world > 1:
            import torch.distributed as dist
            if self.emul:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=self.torch.device("cuda", self.local))

    def sync(self):
        if not self.emul:
            self.torch.cuda.synchronize()

    def timed_ms(self, fn):
        """device time of fn() on torch's current stream (CUDA events); wall time on the emulator"""
        if self.emul:
            t0 = time.perf_counter()
            r = fn()
            return 1000 * (time.perf_counter() - t0), r
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.torch.cuda.synchronize()
        return e0.elapsed_time(e1), r


def run_findclosest(args, rank, world, local):
    """Secondary workloads (BASELINE configs[3] and [4]): greedy rep-DB build at threshold 50 over synthetic
    candidates, then findClosest of synthetic queries against the first --nrep representatives.  Queries are
    sharded by rank against a replicated database; per-query results are all-gathered (12 B/query)."""
    import torch
    import torch.distributed as dist
    import kmers.reps_b200 as kr
    from kmers.reps_b200 import sharding, synth
    dev = Dev(local)
    dev.init_dist(world)
    ncand, nq, nrep = args.ncand, args.nq, args.nrep
    t0 = time.perf_counter()
    res, off, fam = synth.generate(0x5EED0005, ncand + nq, max(1, ncand // 50), xrun_per_10k=10)
    gen_s = time.perf_counter() - t0
    pin = {"auto": 0, "join": kr.FORCE_JOIN, "index": kr.FORCE_INDEX}[args.path]
    ctx = kr.Context(device=local, K=8, flags=kr.IDENT_IS_INFINITY | pin)
    # ---- config 4: greedy build (replicated on every rank: the accept/reject sweep is sequential) ----
    cands = ctx.encode(res=res[: int(off[ncand])], off=off[: ncand + 1])
    db = ctx.db()
    ctx.synchronize()
    ctx.stats_reset()
    t0 = time.perf_counter()
    added = db.greedy_build(cands, 50, batch=args.build_batch)
    build_s = time.perf_counter() - t0
    st_build = ctx.stats()
    nreps_built = int(added.sum())
    acc = np.nonzero(added)[0][:nrep].astype(np.uint32)
    qdb = ctx.db()
    qdb.append_set(cands, acc)
    R = qdb.size
    # ---- config 5: findClosest, queries sharded by rank --------------------------------------------
    lo, hi = sharding.shard_range(nq, rank, world)
    q_res = torch.from_numpy(res[int(off[ncand + lo]): int(off[ncand + hi])]).to(dev.name)
    q_off = off[ncand + lo: ncand + hi + 1] - off[ncand + lo]
    myq = hi - lo
    d_sim = torch.empty(myq, dtype=torch.int32, device=dev.name)
    d_first = torch.empty(myq, dtype=torch.int32, device=dev.name)
    d_last = torch.empty(myq, dtype=torch.int32, device=dev.name)
    chunk = args.q_chunk

    def step():
        for a in range(0, myq, chunk):
            b = min(myq, a + chunk)
            qs = ctx.encode_device(q_res.data_ptr() + int(q_off[a]), q_off[a: b + 1] - q_off[a])
            qdb.find_closest_device(qs, d_sim.data_ptr() + 4 * a, d_first.data_ptr() + 4 * a, d_last.data_ptr() + 4 * a)
            qs.free()        # stream-ordered

    for _ in range(max(1, args.warmup)):
        step()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    dev.sync()
    ctx.stats_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timer_begin()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_end()
    clocks = sampler.stop() if rank == 0 else None
    st = ctx.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev.name)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    gather_ms = None
    if world > 1:
        local_res = torch.stack([d_sim, d_first, d_last], 1)
        gather_ms, allres = dev.timed_ms(lambda: sharding.allgather_ranges(local_res, nq, dist))
        assert allres.shape[0] == nq
    # verify a sample against the oracle
    from oracle import kr_oracle as ko
    o = ko.get()
    cores = os.cpu_count() or 1
    reps_b = o.batch([res[int(off[i]): int(off[i + 1])].tobytes().decode() for i in acc], K=8)
    pick = np.random.default_rng(5 + rank).choice(myq, min(100, myq), replace=False)
    qsel = o.batch([res[int(off[ncand + lo + i]): int(off[ncand + lo + i + 1])].tobytes().decode() for i in pick], K=8)
    bs, bi = o.find_closest(reps_b, qsel, threads=cores)
    tp = torch.from_numpy(pick).to(dev.name)
    got_sim = d_sim[tp].cpu().numpy()
    got_first = d_first[tp].cpu().numpy()
    ok = (got_sim == bs).all() and (got_first == np.where(bs > 0, bi, -1)).all()
    if not ok:
        raise SystemExit("bench: findClosest differs from the oracle on sampled queries")
    used_index = st["scan_launches"] > 0
    # ---- roofline of the scan kernel (per query n_q (8 + 16) + 4 M_q + 8 bytes) --------------------
    # n_q is exact (the sampled queries' set sizes scaled to all queries of the rank); M_q = sum over the representatives of
    # sim(q, rep) = postings the query's keys hold in the index, emitted by the oracle for the verified sample
    hbm, hbm_src = peaks()
    raw = o.sim_matrix(qsel, reps_b, ident_inf=False, threads=cores)
    M_mean = float(raw.sum(axis=1).mean())
    n_mean = float(qsel.sizes().mean())
    roofline = None
    if used_index:
        per_q = n_mean * 24 + 4 * M_mean + 8
        kern_ms = st["scan_ms"] / max(1, st["scan_launches"])
        q_per_launch = myq * args.steps / max(1, st["scan_launches"])
        alg = per_q * q_per_launch
        achieved = alg / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                    "kernel": "k_ix_query<ARGMAX>", "kernel_ms": kern_ms, "kernel_share_of_step": st["scan_ms"] / args.steps / ms_per_step if ms_per_step else None,
                    "peak_source": hbm_src, "algorithmic_bytes": alg, "queries_per_launch": q_per_launch,
                    "bytes_per_query": per_q, "n_q_mean": n_mean, "M_q_mean": M_mean,
                    "note": "SURVEY.md 8d: n_q (8 + 16) + 4 M_q + 8 bytes per query; n_q and M_q are the means over the %d oracle-verified "
                            "sample queries of this rank (M_q = sum of raw similarities to all %d representatives, from the oracle)" % (len(pick), R)}
    cpu = cpu_build = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the reference's algorithm (oracle port: every representative's HashSet probed with the query's k-mers) on a
        # bounded sample of the same queries, all host threads
        ns = min(myq, args.cpu_queries)
        cq = o.batch([res[int(off[ncand + i]): int(off[ncand + i + 1])].tobytes().decode() for i in range(ns)], K=8)
        t0 = time.perf_counter()
        o.find_closest(reps_b, cq, threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": "the first %d queries of the workload against the same %d representatives (%.1f s)" % (ns, R, dt)}
        # the greedy build of the first --cpu-build candidates: the reference's loop is sequential (RepGenomeDbProcessor.java:134-147),
        # so is the port: one thread.  The GPU library builds the same prefix for a like-for-like pair of numbers.
        nb = min(ncand, args.cpu_build)
        if nb:
            cb = o.batch(res=res[: int(off[nb])], off=off[: nb + 1], K=8)
            t0 = time.perf_counter()
            n_cpu, added_cpu, pairs_cpu = o.greedy_build(cb, 50)
            dt = time.perf_counter() - t0
            sub = cands.gather(np.arange(nb, dtype=np.uint32))
            db2 = ctx.db()
            ctx.synchronize()
            t0 = time.perf_counter()
            added_gpu = db2.greedy_build(sub, 50, batch=args.build_batch)
            dt_gpu = time.perf_counter() - t0
            if not (added_gpu == added_cpu).all():
                raise SystemExit("bench: greedy build differs from the oracle on the first %d candidates" % nb)
            db2.free(); sub.free()
            cpu_build = {"value": nb / dt, "unit": "candidates/s", "cores": 1, "kind": "port",
                         "sample": "greedy build (threshold 50) of the first %d candidates: %d representatives, %d pair evaluations with early exit, "
                                   "%.1f s on one thread (the reference's loop is sequential)" % (nb, n_cpu, pairs_cpu, dt),
                         "gpu_same_prefix": {"seconds": dt_gpu, "candidates_per_s": nb / dt_gpu, "identical_result": True}}
    if rank == 0:
        stage_keys = ("encode_ms", "index_ms", "scan_ms", "blocklist_ms", "join_ms", "other_ms")
        line = {
            "metric": "findClosest queries/sec", "value": nq / (ms_per_step / 1e3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "findClosest of %d synthetic queries against a %d-rep DB (BASELINE configs[4]); DB = first %d reps of a "
                                   "greedy build at threshold 50 over %d synthetic candidates (configs[3])" % (nq, R, R, ncand),
                       "K": 8, "path": ("sorted index + warp scan" if used_index else "tile join") + (" (pinned)" if args.path != "auto" else " (library's choice)"),
                       "query_chunk": chunk,
                       "sharding": "queries split over %d rank(s), DB replicated, results all-gathered" % world},
            "pair_equivalents_per_s": nq * R / (ms_per_step / 1e3),
            "roofline": roofline,
            "greedy_build": {"candidates": ncand, "reps": nreps_built, "seconds": build_s, "candidates_per_s": ncand / build_s,
                             "pair_equivalents": int(st_build["pairs"]), "kernel_ms": {k: st_build[k] for k in stage_keys},
                             "launches": int(st_build["kernel_launches"]), "index_builds": int(st_build["index_builds"])},
            "stages_ms_per_step": {k: st[k] / args.steps for k in stage_keys},
            "gpu_launches": int(st["kernel_launches"]), "clocks": clocks, "result_allgather_ms": gather_ms,
            "cpu_baseline": cpu, "cpu_baseline_build": cpu_build,
            "generate_s": gen_s, "verified": "%d sampled queries per rank equal the oracle" % len(pick),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="allvsall", choices=["allvsall", "findclosest"])
    ap.add_argument("--ncand", type=int, default=1000000)
    ap.add_argument("--nq", type=int, default=10000000)
    ap.add_argument("--nrep", type=int, default=50000)
    ap.add_argument("--q-chunk", type=int, default=1000000)
    ap.add_argument("--build-batch", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nseq", type=int, default=100000)
    ap.add_argument("--ref-sample", type=int, default=10000)
    ap.add_argument("--cpu-sample", type=int, default=16000)
    ap.add_argument("--cpu-queries", type=int, default=1000, help="findclosest workload: queries timed on the CPU port")
    ap.add_argument("--cpu-build", type=int, default=10000, help="findclosest workload: candidates of the greedy build timed on the CPU port (BASELINE.md names the first 20,000: "
                         "about 80 s on one core, the loop is sequential; the default keeps the run within its time bound)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--path", default="auto", choices=["auto", "join", "index"],
                    help="device path: the library's own choice, the tile join pinned, or the sorted index + scan pinned")
    ap.add_argument("--elem", default="u16", choices=["u16", "i32"])
    ap.add_argument("--sharded-encode", action="store_true",
                    help="N>1 with --path join: each rank encodes 1/N of the sequences and the block lists are all-gathered (NCCL) and "
                         "imported; the default is that every rank encodes all sequences (no data-path collective)")
    ap.add_argument("--gather-cells", action="store_true",
                    help="N>1: after the timed region, all-gather every rank's sparse result cells (NCCL) and report its cost")
    args = ap.parse_args()
    ensure_built()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "findclosest":
        run_findclosest(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    import kmers.reps_b200 as kr
    from kmers.reps_b200 import sharding

    dev = Dev(local)
    dev.init_dist(world)
    res, off, nfam = workload(args.nseq)
    N = args.nseq
    pin = {"auto": 0, "join": kr.FORCE_JOIN, "index": kr.FORCE_INDEX}[args.path]
    ctx = kr.Context(device=local, K=8, flags=kr.IDENT_IS_INFINITY | pin)
    elem = kr.ELEM_U16 if args.elem == "u16" else kr.ELEM_I32
    esz = 2 if elem == kr.ELEM_U16 else 4
    layout = kr.LAYOUT_BLOCKTRI
    r0, r1, my_elems, _ = sharding.my_stripe(N, layout, rank, world)
    d_res = torch.from_numpy(res).to(dev.name)
    d_out = torch.empty(my_elems * esz + 16, dtype=torch.uint8, device=dev.name)
    total_pairs = N * (N - 1) // 2
    sharded_encode = world > 1 and args.sharded_encode and args.path == "join"

    def step():
        s = sharding.encode_sharded(ctx, d_res, off, dist) if sharded_encode else ctx.encode_device(d_res.data_ptr(), off)
        ctx.sim_matrix_device(s, None, d_out.data_ptr(), elem=elem, layout=layout, row_begin=r0, row_end=r1)
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        dev.sync()

    prev = None
    for _ in range(args.warmup):
        cur = step()
        if prev is not None:
            prev.free()
        prev = cur
    ctx.synchronize()
    if sharded_encode:   # a set imported from block lists has no per-sequence arrays: encode once more, plainly
        tmp = ctx.encode_device(d_res.data_ptr(), off)
        sizes = tmp.sizes().astype(np.int64)
        tmp.free()
    else:
        sizes = prev.sizes().astype(np.int64)

    # ---- timed region: device-resident inputs --------------------------------------------------
    barrier()
    ctx.stats_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timer_begin()
    for _ in range(args.steps):
        cur = step()
        prev.free()          # stream-ordered: the previous step's set goes back to the pool, as a real caller would
        prev = cur
    ms = ctx.timer_end()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    st = ctx.stats()
    prev.free()
    t = torch.tensor([ms], dtype=torch.float64, device=dev.name)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = total_pairs / (ms_per_step / 1e3)
    used_index = st["scan_launches"] > 0

    # ---- check a sample of this rank's cells against the oracle (outside the timed region) -----
    from oracle import kr_oracle as ko
    o = ko.get()
    rng = np.random.default_rng(1234 + rank)
    ii = rng.integers(r0, max(r1, r0 + 1), 400)
    jj = np.array([rng.integers(i, N) for i in ii])
    # half of the sample from inside the row's own family neighbourhood (non-zero cells), half anywhere
    jj[:200] = np.minimum(N - 1, ii[:200] + rng.integers(0, 60, 200))
    host_cells = []
    flat_out = d_out[: my_elems * esz].view(torch.int16 if esz == 2 else torch.int32)
    origin = kr.matrix_elems(N, N, layout, 0, r0)
    for i, j in zip(ii, jj):
        I = i // 256
        idx = kr.matrix_elems(N, N, layout, 0, int(I * 256)) + (i - I * 256) * (N - I * 256) + (j - I * 256) - origin
        host_cells.append(int(flat_out[idx].item()) & 0xFFFFFFFF if esz == 4 else int(flat_out[idx].item()) & 0xFFFF)
    want = [o.similarity(res[int(off[i]):int(off[i + 1])].tobytes().decode(), res[int(off[j]):int(off[j + 1])].tobytes().decode(), 8, True)
            for i, j in zip(ii, jj)]
    verified = host_cells == want
    if not verified and not os.environ.get("KR_BENCH_NOVERIFY"):
        raise SystemExit("bench: GPU result differs from the oracle on sampled cells")
    # size-independent checks of the whole stripe: every diagonal cell is INFINITY, and the number of non-zero cells
    # equals the diagonal + twice-counted in-block pairs + the sparse cell count (checked in tests at full size)
    diag = torch.arange(r0, r1, device=dev.name, dtype=torch.int64)
    DI = diag // 256
    didx = 256 * (DI * N - 128 * DI * (DI - 1)) + (diag - 256 * DI) * (N - 256 * DI) + (diag - 256 * DI) - origin
    if not bool((flat_out[didx].to(torch.int32) & 0xFFFF == kr.INFINITY).all()):
        raise SystemExit("bench: a diagonal cell is not INFINITY")

    # ---- roofline of the dominant kernel ----------------------------------------------------
    hbm, hbm_src = peaks()
    ssum = int(sizes.sum())
    out_bytes = my_elems * esz
    stage_ms = {k: st[k] / args.steps for k in ("encode_ms", "index_ms", "fill_ms", "scan_ms", "blocklist_ms", "join_ms", "other_ms")}
    rows = np.arange(r0, r1)
    roofline_scan = None
    if used_index:
        # The index path leaves a dense stripe as zeros + the cells of pairs that share a k-mer.  When the call builds its index
        # (every step here: the batch is re-encoded), the zeros come from k_zero_fill on a second stream BESIDE the index build
        # and the scan (k_ix_query<DENSE>) only writes the non-zero cells; otherwise the scan writes the zeros row by row itself.
        # The dominant kernel is whichever of the two is longer; the other one is reported as roofline_scan / in stages.
        frac_rows = (r1 - r0) / max(1, N)
        walked = int(st["index_walk"] * frac_rows)
        prezeroed = st["fill_launches"] > 0
        scan_ms = st["scan_ms"] / max(1, st["scan_launches"])
        fill_ms = st["fill_ms"] / max(1, st["fill_launches"])
        # SURVEY.md 8d, index-served paths, per query row: n_q (8 + 16) + 4 M_q (keys, one 16-byte index probe per key, 4 B per
        # posting walked; M = the index's own statistic, sum over posting lists of m^2) (+ the row's cells when the scan writes them)
        scan_alg = int(sizes[rows].sum()) * 24 + 4 * walked + (0 if prezeroed else out_bytes)
        scan_note = ("per row: n (8 + 16) B of keys and index probes + 4 B per posting walked (sum of m^2 over the posting lists = %d for the "
                     "whole batch)" % st["index_walk"]) + ("" if prezeroed else " + the row's result cells written once (zeros included)")
        if prezeroed and fill_ms >= scan_ms:
            kern, kern_ms, alg = "k_zero_fill", fill_ms, out_bytes
            alg_note = ("the rank's dense result stripe written once (it is zeros except for the cells the scan adds afterwards); the kernel runs "
                        "on a second stream beside the index build (index_ms in stages), so its time includes the contention with it")
            roofline_scan = {"kernel": "k_ix_query<DENSE>", "kernel_ms": scan_ms, "algorithmic_bytes": scan_alg,
                             "achieved": scan_alg / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0, "unit": "GB/s", "note": scan_note}
        else:
            kern, kern_ms, alg, alg_note = "k_ix_query<DENSE>", scan_ms, scan_alg, scan_note
    else:
        # tile join (k_join<MATRIX>): compulsory DRAM traffic of one launch: key lists once + result cells once
        kern, kern_ms = "k_join<MATRIX>", st["join_ms"] / max(1, st["join_launches"])
        alg = ssum * 8 + out_bytes
        alg_note = "compulsory DRAM traffic of one launch (SURVEY.md 8d.3): the batch's key lists once (8 B per distinct k-mer) + this rank's result cells once"
    achieved = alg / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                "kernel": kern, "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_per_step if ms_per_step else None,
                "peak_source": hbm_src, "algorithmic_bytes": alg, "note": alg_note,
                "traffic_note": "ncu dram bytes are not measured inside bench.py; see profiles/ for captures (none exist for this revision "
                                "if GPU access was closed)"}
    # whole step against the compulsory traffic: residues in + key lists written and read + result written
    step_bytes = int(len(res)) + 2 * ssum * 8 + out_bytes
    step_roof = {"bytes": step_bytes, "achieved": step_bytes / (ms_per_step / 1e3) / 1e9, "frac_of_hbm_peak": step_bytes / (ms_per_step / 1e3) / 1e9 / hbm,
                 "note": "whole step (encode + index + scan, max over ranks) against its compulsory bytes: residues + key lists out and in + result"}
    # the per-pair streaming figure of SURVEY.md 8d.1 (8(|A|+|B|)+4 B per pair): a WORK-EQUIVALENCE number, not a bound for an
    # algorithm that never streams two key lists per pair
    suffix = np.concatenate([np.cumsum(sizes[::-1])[::-1], [0]])
    key_stream = int((8 * (sizes[rows] * (N - 1 - rows) + suffix[rows + 1]) + 4 * (N - 1 - rows)).sum())
    contract = {"algorithmic_bytes": key_stream, "achieved": key_stream / (ms_per_step / 1e3) / 1e9, "unit": "GB/s",
                "frac_of_hbm_peak": key_stream / (ms_per_step / 1e3) / 1e9 / hbm,
                "note": "SURVEY.md 8d.1 pairwise key-stream figure over the whole step; not a roofline for an index scan or a block-level "
                        "join (reported because north_star quotes it)"}

    # ---- e2e: host buffers through the C ABI -----------------------------------------------------
    e2e = e2e_dense = None
    if not args.no_e2e:
        h_res = ctx.host_alloc(max(len(res), 1))
        h_res[: len(res)] = res
        cap = 1 << 23
        h_cells = ctx.host_alloc(cap * 8, dtype=np.uint64)          # packed cells: row << 40 | col << 16 | sim
        ncell = [0]

        def timed(fn):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            dev.sync()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev.name)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        def sparse_step():
            s = ctx.encode(res=h_res[: len(res)], off=off)
            ncell[0] = len(ctx.sim_matrix_sparse64(s, None, min_sim=1, layout=kr.LAYOUT_TRIANGLE, row_begin=r0, row_end=r1, out=h_cells))
            s.free()
        dt = timed(sparse_step)
        e2e = {"value": total_pairs / (dt / args.e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(len(res) + 16 * (N + 1)),
               "d2h_bytes_per_step": int(8 + 8 * ncell[0]), "ms_per_step": 1000 * dt / args.e2e_steps, "steps": args.e2e_steps,
               "cells_per_step": ncell[0],
               "api": "kr_encode + kr_sim_matrix_sparse64 (pinned host buffers): the cells with sim >= 1 of this rank's rows, 8 bytes each; "
                      "every other cell is 0"}
        # the e2e result is checked too: every sampled non-zero cell of the device-resident run above is among the cells, with its value
        rr, cc, vv = kr.unpack_cells(h_cells[: ncell[0]])
        got_cells = dict(zip(zip(rr.tolist(), cc.tolist()), vv.tolist()))
        for (i, j), w in zip(zip(ii.tolist(), jj.tolist()), want):
            if i < j and w > 0 and got_cells.get((i, j)) != w:
                raise SystemExit("bench: e2e sparse result differs from the oracle at (%d, %d)" % (i, j))
            if i < j and w == 0 and (i, j) in got_cells:
                raise SystemExit("bench: e2e sparse result has a cell the oracle does not (%d, %d)" % (i, j))
        del h_cells
        # the dense variant needs a pinned host buffer as large as the stripe (10 GB at N = 1): informational, and a box that
        # cannot pin that much must not cost the run its line (every rank takes the same decision: the size is the same test)
        try:
            h_out = ctx.host_alloc(my_elems * esz, dtype=np.uint16 if esz == 2 else np.int32)
        except kr.KrError as e:
            h_out = None
            e2e_dense = {"skipped": "pinned host allocation of %d bytes failed: %s" % (my_elems * esz, e)}
        ok_all = torch.tensor([1 if h_out is not None else 0], dtype=torch.int32, device=dev.name)
        if world > 1:
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if int(ok_all.item()) == 1:
            def dense_step():
                s = ctx.encode(res=h_res[: len(res)], off=off)
                ctx.sim_matrix(s, None, elem=elem, layout=layout, row_begin=r0, row_end=r1, out=h_out)
                s.free()
            dt = timed(dense_step)
            e2e_dense = {"value": total_pairs / (dt / args.e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(len(res) + 16 * (N + 1)),
                         "d2h_bytes_per_step": int(my_elems * esz), "ms_per_step": 1000 * dt / args.e2e_steps, "steps": args.e2e_steps,
                         "api": "kr_encode + kr_sim_matrix (pinned host buffers): the dense stripe over PCIe"}
        elif e2e_dense is None:
            e2e_dense = {"skipped": "another rank could not pin its host buffer"}

    gather = None
    if args.gather_cells and world > 1:
        cap = 1 << 22
        d_cells = torch.zeros((cap, 4), dtype=torch.int32, device=dev.name)
        dev.sync()                                              # torch's fill is done before the library's stream writes the buffer
        s = ctx.encode_device(d_res.data_ptr(), off)
        cnt = ctx.sim_matrix_sparse_device(s, None, d_cells.data_ptr(), cap, min_sim=1, layout=kr.LAYOUT_TRIANGLE, row_begin=r0, row_end=r1)
        sharding.allgather_cells(d_cells, cnt, cap, dist)      # warm-up
        barrier()
        g_ms, allc = dev.timed_ms(lambda: sharding.allgather_cells(d_cells, cnt, cap, dist))
        tg = torch.tensor([g_ms], dtype=torch.float64, device=dev.name)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"ms": float(tg.item()), "cells_total": int(allc.shape[0]), "bytes_per_rank": int(cap * 16)}
        s.free()

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            v, pairs, dt, n = cpu_baseline(res, off, args.cpu_sample, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "all i<j pairs of the first %d sequences of the workload (%d pairs, %.1f s)" % (n, pairs, dt)}
            # the reference's own loops are single-threaded (DistanceMatrixProcessor.java:71-76): the same port on ONE thread,
            # on a smaller sample (2,000 x 2,000)
            v1, pairs1, dt1, n1 = cpu_baseline(res, off, min(2000, args.cpu_sample), 1)
            cpu["single_thread"] = {"value": v1, "unit": UNIT, "cores": 1,
                                    "sample": "all i<j pairs of the first %d sequences (%d pairs, %.1f s)" % (n1, pairs1, dt1)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "all-vs-all 8-mer similarity, %d synthetic ~300-aa seed proteins, %d families (BASELINE configs[2])" % (N, nfam),
                       "K": 8, "nseq": N, "pairs": total_pairs, "result": "%s block-triangle, %d bytes on this rank" % (args.elem, my_elems * esz),
                       "path": ("sorted index + warp scan" if used_index else "tile join") + (" (pinned)" if args.path != "auto" else " (library's choice)"),
                       "sharding": ("256-row stripes of the result over %d rank(s); " % world) +
                                   ("each rank encodes 1/%d of the sequences, block lists all-gathered (NCCL)" % world if sharded_encode
                                    else "every rank encodes and indexes the whole batch, no data-path collective"),
                       "cache": "inputs+outputs larger than L2 (key lists %d MB, result %d MB per rank)" % (ssum * 8 >> 20, my_elems * esz >> 20)},
            "roofline": roofline, "roofline_scan": roofline_scan, "roofline_step": step_roof, "roofline_contract_8d1": contract, "cpu_baseline": cpu, "e2e": e2e, "e2e_dense": e2e_dense,
            "gpu_launches": int(st["kernel_launches"]), "clocks": clocks,
            "stages_ms_per_step": stage_ms, "index": {"entries": int(st["index_entries"]), "postings_walked_by_a_self_scan": int(st["index_walk"])},
            "verified": "400 sampled cells per rank equal the oracle (200 of them near the diagonal, inside families); every diagonal cell is INFINITY",
        }
        if gather:
            line["cells_allgather"] = gather
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
