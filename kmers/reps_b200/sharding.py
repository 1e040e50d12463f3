"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the path shards and what is exchanged.

The path has no data-path collective: all-vs-all shards the 256-row stripes of the result (every rank encodes and
indexes the whole batch, then scans only its rows), findClosest shards the queries against a replicated database.
The only exchanges are result gathers, done here with plain torch.distributed collectives (NCCL on GPUs; gloo in the
CPU tests).  encode_sharded() is the optional block-list exchange of the tile-join path."""
from __future__ import annotations

import numpy as np

from . import binding as kr


def my_stripe(n: int, layout: int, rank: int, world: int):
    """rows [r0, r1) of an n x n self matrix owned by `rank`, their cell count and the offset of the first cell
    in the whole matrix (cells of lower-ranked stripes)"""
    rb = kr.matrix_stripes(n, n, layout, world)
    r0, r1 = int(rb[rank]), int(rb[rank + 1])
    return r0, r1, kr.matrix_elems(n, n, layout, r0, r1), kr.matrix_elems(n, n, layout, 0, r0)


def stripe_table(n: int, layout: int, world: int):
    rb = kr.matrix_stripes(n, n, layout, world)
    counts = [kr.matrix_elems(n, n, layout, int(rb[k]), int(rb[k + 1])) for k in range(world)]
    return rb, counts


def allgather_stripes(local, n: int, layout: int, dist, out=None):
    """Every rank ends up with the whole matrix (1-D, `layout` order).  `local` is this rank's stripe as a 1-D
    torch tensor.  Stripes differ slightly in size (block granularity), so each rank broadcasts its own into its
    slice of the result."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    rb, counts = stripe_table(n, layout, world)
    total = int(sum(counts))
    if out is None:
        out = torch.empty(total, dtype=local.dtype, device=local.device)
    pos = 0
    for k in range(world):
        sl = out[pos:pos + counts[k]]
        if k == rank:
            sl.copy_(local[: counts[k]])
        if counts[k]:
            dist.broadcast(sl.view(torch.uint8), src=k)   # bytes: NCCL has no 16-bit integer type
        pos += counts[k]
    return out


def shard_range(n: int, rank: int, world: int):
    """contiguous query range of `rank`: sizes differ by at most one"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_ranges(local, n: int, dist):
    """gather per-query results sharded by shard_range back into query order on every rank"""
    import torch
    world = dist.get_world_size()
    base, extra = divmod(n, world)
    width = base + (1 if extra else 0)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = []
    for k in range(world):
        lo, hi = shard_range(n, k, world)
        out.append(parts[k][: hi - lo])
    return torch.cat(out, 0)


def blocktri_from_full(full: np.ndarray) -> np.ndarray:
    """KR_LAYOUT_BLOCKTRI cells of a full n x n matrix (row i keeps columns >= 256*(i//256))"""
    n = full.shape[0]
    return np.concatenate([full[i, 256 * (i // 256):] for i in range(n)]) if n else full.reshape(0)


# ---- sharded encode: every rank encodes a slice, block lists are all-gathered (NCCL) ----------------------
class _DevPtr:
    """zero-copy view of library-owned device memory for torch (CUDA array interface)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr: int, nbytes: int, device):
    """library-owned memory as a uint8 torch tensor on `device` (CUDA; or CPU when the library is the test emulator)"""
    import torch
    if torch.device(device).type == "cuda":
        return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)
    import ctypes
    return torch.frombuffer((ctypes.c_uint8 * int(nbytes)).from_address(int(ptr)), dtype=torch.uint8)


def _sync(device):
    import torch
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()


def block_shards(nseq: int, world: int):
    """contiguous sequence ranges, one per rank, cut at multiples of 256 sequences"""
    nblk = (nseq + kr.TILE - 1) // kr.TILE
    return [min(nseq, kr.TILE * ((nblk * k) // world)) for k in range(world)] + [nseq]


def encode_sharded(ctx, d_res, off, dist):
    """Encode the whole batch cooperatively: this rank encodes and block-lists its slice, the slices' block
    lists are exchanged with ONE all_gather, and the result is imported as one set on every rank.
    d_res: torch uint8 CUDA tensor with ALL residues (replicated, 1 B/residue); off: host offsets.
    Every rank lays its six arrays out in a segment whose field sizes are capacity bounds all ranks compute
    from the offsets alone, so no size exchange is needed."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    nseq = len(off) - 1
    cuts = block_shards(nseq, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    local = ctx.encode_device(d_res.data_ptr(), off[lo: hi + 1])
    bl = ctx.export_blocklists(local)                      # synchronises: the arrays are complete
    shards = []
    for k in range(world):
        b = kr.BlockLists()
        b.nseq = cuts[k + 1] - cuts[k]
        b.nblk = (b.nseq + kr.TILE - 1) // kr.TILE
        b.entries = ctx.blocklist_capacity(off[cuts[k]: cuts[k + 1] + 1])
        b.lo_size, b.nchunk = bl.lo_size, bl.nchunk
        shards.append(b)
    names = ("hi", "lo", "chunk_off", "chunk_cnt", "blk_base", "meta")
    field, width = {}, 0
    for name in names:                                     # common segment layout: field sizes = max over ranks
        field[name] = width
        width += (max(shards[k].arrays()[name][1] for k in range(world)) + 255) // 256 * 256
    send = torch.empty(width, dtype=torch.uint8, device=d_res.device)
    mine = shards[rank]
    for name in names:
        nbytes = mine.arrays()[name][1]
        if nbytes:
            send[field[name]: field[name] + nbytes].copy_(_view(getattr(bl, name), nbytes, d_res.device))
    recv = torch.empty(world * width, dtype=torch.uint8, device=d_res.device)
    dist.all_gather_into_tensor(recv, send)
    for k in range(world):
        for name in names:
            setattr(shards[k], name, recv.data_ptr() + k * width + field[name])
    _sync(d_res.device)                                    # the gathered arrays are complete before the library reads them
    full = ctx.import_blocklists(shards, d_res.data_ptr(), off)
    full._keep = (recv, local)   # released with the set: by then the import's copies have long completed
    return full


def allgather_cells(local_cells, count: int, cap: int, dist):
    """Sparse result cells of every rank on every rank: ONE all_gather of capacity-padded (cap, 4) uint32 blocks plus
    one of the counts (north_star: "NCCL allgather of per-shard results").  Returns the (total, 4) tensor."""
    import torch
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=local_cells.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([count], dtype=torch.int64, device=local_cells.device))
    recv = torch.empty((world * cap, 4), dtype=local_cells.dtype, device=local_cells.device)
    dist.all_gather_into_tensor(recv, local_cells[:cap].contiguous())
    return torch.cat([recv[k * cap: k * cap + int(counts[k].item())] for k in range(world)], 0)


# ---- greedy representative-set build with the candidate-vs-database scoring split over the ranks ------------
def greedy_build_sharded(n_cands: int, threshold: int, batch: int, dist, score_slice, survivor_matrix, accept, device=None):
    """RepGenomeDb.checkGenome over `n_cands` candidates in order (RepGenomeDbProcessor.java:160-163,
    BaseGenomeProcessor.java:203-204), every rank holding a replica of the database.

    Per batch of `batch` candidates: each rank scores its shard_range of the batch against its replica, the best
    similarities are all-gathered (4 B per candidate), and every rank then runs the same deterministic in-order
    sweep over the survivors (candidates no current representative covers), so the replicas stay identical without
    a broadcast.  The sweep's accept/reject chain is sequential by nature: only the scoring scales with the ranks.

        score_slice(lo, hi)      -> int array: best similarity of candidates [lo, hi) to the current database
        survivor_matrix(idx)     -> (len(idx) x len(idx)) similarities among candidates idx (ascending)
        accept(idx)              -> append candidates idx to this rank's replica
        device                   -> where the exchanged tensor lives (a CUDA device for NCCL; None = CPU for gloo)
    Returns the uint8 `added` flags, identical on every rank."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    added = np.zeros(n_cands, dtype=np.uint8)
    db_size = 0
    for pos in range(0, n_cands, batch):
        n = min(batch, n_cands - pos)
        if db_size:
            lo, hi = shard_range(n, rank, world)
            mine = np.asarray(score_slice(pos + lo, pos + hi), dtype=np.int32).reshape(hi - lo)
            t = torch.from_numpy(mine)
            best = allgather_ranges(t if device is None else t.to(device), n, dist).cpu().numpy()
            surv = pos + np.nonzero(best < threshold)[0]
        else:
            surv = pos + np.arange(n)
        if len(surv) == 0:
            continue
        acc = []
        if len(surv) == 1:
            acc = [0]
        else:
            m = np.asarray(survivor_matrix(surv))
            for s in range(len(surv)):                       # checkGenome, in candidate order
                if not acc or not (m[s, acc] >= threshold).any():
                    acc.append(s)
        idx = surv[np.asarray(acc, dtype=np.int64)]
        added[idx] = 1
        accept(idx.astype(np.uint32))
        db_size += len(idx)
    return added


def greedy_build_gpu(ctx, db, cands, threshold: int, dist, batch: int = 4096, device=None):
    """greedy_build_sharded on this rank's GPU: `db` is this rank's replica (kr_db), `cands` the encoded candidates
    (the same set on every rank)."""
    def score_slice(lo, hi):
        if hi <= lo:
            return np.zeros(0, np.int32)
        return db.find_closest(cands.gather(np.arange(lo, hi, dtype=np.uint32)))[0]

    def survivor_matrix(idx):
        return ctx.sim_matrix(cands.gather(idx.astype(np.uint32)), elem=kr.ELEM_U16)

    return greedy_build_sharded(cands.count, threshold, batch, dist, score_slice, survivor_matrix, lambda idx: db.append_set(cands, idx), device=device)
