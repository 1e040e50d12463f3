"""TEST INFRASTRUCTURE: the N > 1 paths end to end on CPU -- gloo ranks, each with its own context of the emulator
library, run kmers/reps_b200/sharding.py exactly as the GPU ranks do (torch CPU tensors stand in for device memory).

    python tests/emul/run_emul_sharded.py <world>
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kmers.reps_b200 as kr
        from kmers.reps_b200 import binding, sharding, synth
        from oracle import kr_oracle as ko
        assert "libkmersreps_emul" in kr.lib_path()
        o = ko.get()
        res, off, _ = synth.generate(41, 900, 10)
        seqs = [s[:110] for s in synth.to_strings(res, off)]
        seqs[3] = seqs[700]                                      # equal sequences encoded by different ranks
        res, off = kr.flat(seqs)
        n = len(seqs)
        want = sharding.blocktri_from_full(o.sim_matrix(o.batch(seqs, K=8), ident_inf=True, threads=2))
        d_res = torch.from_numpy(res)
        layout = kr.LAYOUT_BLOCKTRI
        r0, r1, elems, first = sharding.my_stripe(n, layout, rank, world)
        for mode in ("blocklists", "replicated-join", "replicated-index"):
            flags = kr.IDENT_IS_INFINITY | (kr.FORCE_INDEX if mode == "replicated-index" else kr.FORCE_JOIN)
            ctx = kr.Context(device=0, K=8, flags=flags)
            s = sharding.encode_sharded(ctx, d_res, off, dist) if mode == "blocklists" else ctx.encode_device(d_res.data_ptr(), off)
            out = torch.zeros(elems + 8, dtype=torch.int16)
            ctx.sim_matrix_device(s, None, out.data_ptr(), elem=kr.ELEM_U16, layout=layout, row_begin=r0, row_end=r1)
            ctx.synchronize()
            got = out[:elems].numpy().view(np.uint16)
            assert (got == want[first:first + elems]).all(), (mode, rank)
            whole = sharding.allgather_stripes(out[:elems], n, layout, dist)
            assert (whole.numpy().view(np.uint16) == want).all(), (mode, rank)
            if mode != "blocklists":
                # the sparse cells of this rank's stripe, all-gathered: every rank ends up with every cell (j > i, sim >= 1)
                cap = 1 << 15
                cells = torch.zeros((cap, 4), dtype=torch.int32)
                cnt = ctx.sim_matrix_sparse_device(s, None, cells.data_ptr(), cap, min_sim=1, layout=kr.LAYOUT_TRIANGLE, row_begin=r0, row_end=r1)
                ctx.synchronize()
                allc = sharding.allgather_cells(cells, cnt, cap, dist).numpy()
                full = o.sim_matrix(o.batch(seqs, K=8), ident_inf=True, threads=2)
                ii, jj = np.nonzero(np.triu(full, 1))
                assert {(int(r), int(c), int(v)) for r, c, v, _ in allc} == {(int(i), int(j), int(full[i, j])) for i, j in zip(ii, jj)}, (mode, rank)
            ctx.close()
        q.put((rank, "ok"))
    except Exception as e:
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 150 + world
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(1200)
    res = sorted(q.get(timeout=5) for _ in range(world))
    if res != [(r, "ok") for r in range(world)]:
        print(res)
        sys.exit(1)
    print("OK sharded", world)
