"""Host-side sharding logic (mcvslam_b200/shard.py) on the gloo backend, world_size 2, CPU only. The compute callables are
injected checkers (the oracle) — on GPUs the same code runs with the CUDA engine's callables (tests/test_gpu_shard.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcvslam_b200 import shard, synth


def test_blocks_cover_and_balance():
    for n in (0, 1, 7, 8, 4096, 1048576):
        for w in (1, 2, 4, 8):
            bl = shard.blocks(n, w)
            assert bl[0][0] == 0 and bl[-1][1] == n
            assert all(bl[i][1] == bl[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in bl]
            assert max(sizes) - min(sizes) <= 1


def test_consecutive_pairs_cover_once():
    for n in (1, 2, 5, 4096):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                pb, pe, fb, fe = shard.consecutive_pairs(n, r, w)
                assert fb <= pb and (pe == pb or fe == pe + 1) and fe <= n
                seen += list(range(pb, pe))
            assert seen == list(range(max(0, n - 1)))


def test_merge_top2_is_lexicographic():
    rng = np.random.default_rng(0)
    d = rng.integers(0, 4, (200, 4)); i = np.stack([rng.permutation(50)[:4] for _ in range(200)])
    ia, da = torch.tensor(i[:, :2], dtype=torch.int32), torch.tensor(d[:, :2], dtype=torch.int32)
    ib, db = torch.tensor(i[:, 2:], dtype=torch.int32), torch.tensor(d[:, 2:], dtype=torch.int32)
    ib[:10] = -1; db[:10] = 0x7FFFFFFF
    mi, md = shard.merge_top2(ia, da, ib, db)
    for r in range(200):
        c = sorted((int(dd), int(ii)) for dd, ii in zip(torch.cat([da[r], db[r]]), torch.cat([ia[r], ib[r]])) if ii >= 0)
        assert [(int(md[r, k]), int(mi[r, k])) for k in range(2)] == c[:2]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q_np, t_np, frames, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    def knn2_fn(qb, tt, off):      # checker as the compute callable (tests only)
        if qb.shape[0] == 0:
            return torch.zeros((0, 2), dtype=torch.int32), torch.zeros((0, 2), dtype=torch.int32)
        dm, k = O.knn2_bf(qb.numpy(), tt.numpy())
        idx = torch.tensor(dm["trainIdx"].astype(np.int32)) + off
        dst = torch.tensor(dm["distance"].astype(np.int32))
        return idx, dst

    t = shard.broadcast_descriptors(torch.from_numpy(t_np) if rank == 0 else None, len(t_np), "cpu")
    idx, dst = shard.knn2_sharded(torch.from_numpy(q_np), t, knn2_fn, tile=150)

    def process_fn(block):
        nb = len(block)
        cnt = torch.zeros((nb, 3), dtype=torch.int32); csum = torch.zeros((nb, 3), dtype=torch.int64)
        orb = O.Orb(300, 1.2, 4, 28, 15)
        for f in range(nb):
            for c in range(3):
                n, k, d = orb.extract(block[f, c])
                cnt[f, c] = n; csum[f, c] = int(d.astype(np.int64).sum())
        return dict(counts=cnt, checksum=csum)

    out = shard.process_frames(frames, process_fn)
    if rank == 0:
        ret["idx"] = idx.numpy(); ret["dist"] = dst.numpy(); ret["counts"] = out["counts"].numpy(); ret["checksum"] = out["checksum"].numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process(oracle):
    q = synth.descriptors(301, 1, True); t = synth.descriptors(400, 2, True)
    q[:, 2:] = 0; t[:, 2:] = 0        # heavy ties: the cross-tile / cross-rank merge must keep (distance, index) order
    frames = np.stack([synth.triplet(s, 320, 240) for s in (1, 2, 3)])
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), q, t, frames, ret), nprocs=2, join=True)
    ref, k = oracle.knn2_bf(q, t)
    assert np.array_equal(ret["idx"], ref["trainIdx"]) and np.array_equal(ret["dist"], ref["distance"].astype(np.int32))
    orb = oracle.Orb(300, 1.2, 4, 28, 15)
    for f in range(3):
        for c in range(3):
            n, kp, d = orb.extract(frames[f, c])
            assert ret["counts"][f, c] == n and ret["checksum"][f, c] == int(d.astype(np.int64).sum())
