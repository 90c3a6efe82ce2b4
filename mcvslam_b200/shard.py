"""Multi-GPU sharding of the path: one process per GPU (torchrun), `torch.distributed` for the plumbing.

The path has no data exchange inside the compute: frames are independent units and a brute-force 2-NN is
independent per query, so
  * frame batches (BASELINE configs[1], configs[3]) are cut into contiguous blocks of frames per rank, every rank runs the
    whole extract + stereo (+ consecutive-frame match) pipeline on its block with its own engine handle, and the ONLY
    collective is the gather of the results at the end (all_gather of fixed-capacity records — NCCL over NVLink on GPUs);
  * the large-scale match (configs[4]) shards the QUERIES; every rank holds the full train set (broadcast once), computes
    its rows of the 2-NN table with the device-resident kernel, and the rows are all_gathered. Sharding the train set
    instead would need a lexicographic (distance, index) merge across ranks — deliberately avoided.

Everything here is host logic over torch tensors; the compute callables default to the CUDA engine (mcvslam_b200.api)
and there is no CPU fallback. tests/ drive the same code on the gloo backend with world_size 2 by injecting a checker as
the compute callable.
"""
import numpy as np
import torch
import torch.distributed as dist


def block(n_units, rank, world):
    """Contiguous block [begin, end) of rank `rank`: the first n_units % world ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def blocks(n_units, world):
    return [block(n_units, r, world) for r in range(world)]


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_ragged(local, n_units, dim0_per_unit=1):
    """Gathers per-rank tensors that hold `block(n_units, rank, world)` units along dim 0 (each unit = dim0_per_unit rows)
    into the full tensor in unit order on every rank. Ranks pad to the largest block so that one all_gather suffices."""
    rank, world = _world()
    if world == 1:
        return local
    sizes = [(e - b) * dim0_per_unit for b, e in blocks(n_units, world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], 0)


def process_frames(frames, process_fn, gather=True):
    """frames: (n_frames, 3, H, W) u8 array known to every rank (or None on ranks that only need their block — then pass
    the block via process_fn's closure). process_fn(frames_block) -> dict of torch tensors whose dim 0 is the frame
    index within the block. Returns the dict for ALL frames (gather=True) or the local block."""
    rank, world = _world()
    n = len(frames)
    b, e = block(n, rank, world)
    local = process_fn(frames[b:e])
    if not gather or world == 1:
        return local
    return {k: all_gather_ragged(v, n) for k, v in local.items()}


def rig_process_fn(rig, device):
    """process_fn for process_frames backed by the CUDA engine: host block in, device tensors out."""
    from . import api as A

    def fn(block_np):
        nb = len(block_np)
        cap = rig.cap
        if nb == 0:
            z = lambda *s, dt=torch.uint8: torch.zeros(s, dtype=dt, device=device)
            return dict(kps=z(0, 3, cap, 28), desc=z(0, 3, cap, 32), counts=z(0, 3, dt=torch.int32), u_right=z(0, cap, dt=torch.float32),
                        depth_left=z(0, cap, dt=torch.float32))
        out = rig.process(np.ascontiguousarray(block_np))
        return dict(kps=torch.from_numpy(out["kps"].view(np.uint8).reshape(nb, 3, cap, 28)).to(device),
                    desc=torch.from_numpy(out["desc"]).to(device), counts=torch.from_numpy(out["counts"]).to(device),
                    u_right=torch.from_numpy(out["u_right"]).to(device), depth_left=torch.from_numpy(out["depth_left"]).to(device))
    return fn


def consecutive_pairs(n_frames, rank, world):
    """configs[3]: 2-NN between consecutive frames i, i+1. Pair i belongs to the rank that owns frame i; that rank also
    extracts frame i+1 (one halo frame recomputed locally at the block boundary — no exchange). Returns (pair_begin,
    pair_end, frame_begin, frame_end)."""
    b, e = block(n_frames, rank, world)
    pb, pe = b, min(e, n_frames - 1)
    if pe <= pb:
        return b, b, b, e
    return pb, pe, b, pe + 1


def knn2_sharded(q, t, knn2_fn, gather=True, tile=1 << 21):
    """configs[4] (LargeScaleMatching): q (nq, 32) u8 and t (nt, 32) u8 torch tensors, identical on every rank (use
    broadcast_descriptors first when only rank 0 has them). Rank r matches query rows block(nq, r, world) against ALL of t.
    knn2_fn(q_block, t_tile, train_offset) -> (idx (n, 2) int32, dist (n, 2) int32) with (distance, trainIdx)
    lexicographic order, -1 / INT_MAX padding; train tiles of at most `tile` rows are merged here with the same order.
    Returns (idx, dist) for all nq queries (gather=True) or for the local block."""
    rank, world = _world()
    nq, nt = q.shape[0], t.shape[0]
    b, e = block(nq, rank, world)
    qb = q[b:e]
    idx = torch.full((e - b, 2), -1, dtype=torch.int32, device=q.device)
    dst = torch.full((e - b, 2), 0x7FFFFFFF, dtype=torch.int32, device=q.device)
    for t0 in range(0, nt, tile):
        i2, d2 = knn2_fn(qb, t[t0: t0 + tile], t0)
        idx, dst = merge_top2(idx, dst, i2, d2)
    if not gather or world == 1:
        return idx, dst
    return all_gather_ragged(idx, nq), all_gather_ragged(dst, nq)


def merge_top2(idx_a, dist_a, idx_b, dist_b):
    """Lexicographic (distance, index) top-2 of two top-2 tables (n, 2) — the order cv::BFMatcher / the first-party loop
    produce. -1 entries (distance INT_MAX) lose to everything. A four-element min / max network on packed
    (distance << 32 | index) keys: elementwise, no sort."""
    def keys(i, d):
        i = i.to(torch.int64)
        return (d.to(torch.int64) << 32) | torch.where(i < 0, torch.full_like(i, 0xFFFFFFFF), i)
    ka, kb = keys(idx_a, dist_a), keys(idx_b, dist_b)
    a0, a1 = torch.minimum(ka[:, 0], ka[:, 1]), torch.maximum(ka[:, 0], ka[:, 1])
    b0, b1 = torch.minimum(kb[:, 0], kb[:, 1]), torch.maximum(kb[:, 0], kb[:, 1])
    m0 = torch.minimum(a0, b0)
    m1 = torch.minimum(torch.maximum(a0, b0), torch.minimum(a1, b1))
    k = torch.stack([m0, m1], 1)
    i = k & 0xFFFFFFFF
    i = torch.where(i == 0xFFFFFFFF, torch.full_like(i, -1), i)
    return i.to(torch.int32), (k >> 32).to(torch.int32)


def broadcast_descriptors(t, n_rows, device, src=0):
    """Makes the (n_rows, 32) u8 descriptor set of rank `src` resident on every rank (ncclBroadcast over NVLink)."""
    rank, world = _world()
    if rank != src:
        t = torch.empty((n_rows, 32), dtype=torch.uint8, device=device)
    if world > 1:
        dist.broadcast(t, src)
    return t


def engine_knn2_fn(stream=None):
    """knn2_fn backed by the device-resident brute-force kernel (mcv_knn2_bf_device)."""
    from . import api as A
    L = A.lib()

    def fn(qb, tt, train_offset):
        assert qb.is_cuda and tt.is_cuda, "the engine matches device-resident descriptors; there is no CPU fallback"
        qb = qb.contiguous(); tt = tt.contiguous()
        idx = torch.empty((qb.shape[0], 2), dtype=torch.int32, device=qb.device)
        dst = torch.empty((qb.shape[0], 2), dtype=torch.int32, device=qb.device)
        s = stream if stream is not None else torch.cuda.current_stream(qb.device).cuda_stream
        A._check(L.mcv_knn2_bf_device(qb.data_ptr(), qb.shape[0], tt.data_ptr(), tt.shape[0], int(train_offset), idx.data_ptr(), dst.data_ptr(), s))
        return idx, dst
    return fn
