"""Multi-GPU paths of configs[3] / configs[4] on REAL NCCL: two ranks, one B200 each (skipped on a one-GPU box). The same host
logic as tests/test_shard_gloo.py (mcvslam_b200/shard.py), but with the CUDA engine as the compute callable and the collectives
over NCCL: knn2_sharded (+ broadcast_descriptors; tie-heavy, several train tiles, tensor-core and integer-pipe sizes) and
process_frames (full fixed-capacity result records gathered), both against the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch

from mcvslam_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q_np, t_np, frames, out_path):
    import torch.distributed as dist
    from mcvslam_b200 import api as A, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    for name, tile in (("one_tile", 1 << 21), ("tiled", 7000)):
        t = shard.broadcast_descriptors(torch.from_numpy(t_np).to(dev) if rank == 0 else None, len(t_np), dev)
        idx, dst = shard.knn2_sharded(torch.from_numpy(q_np).to(dev), t, shard.engine_knn2_fn(), tile=tile)
        res[name] = (idx.cpu().numpy(), dst.cpu().numpy())
    rig = A.Rig(300, 1.2, 4, 28, 15, device=rank)
    out = shard.process_frames(frames, shard.rig_process_fn(rig, dev))
    torch.cuda.synchronize(dev)
    if rank == 0:
        np.savez(out_path, idx0=res["one_tile"][0], dst0=res["one_tile"][1], idx1=res["tiled"][0], dst1=res["tiled"][1],
                 **{k: v.cpu().numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_world2_nccl_matches_oracle(oracle, tmp_path):
    import torch.multiprocessing as mp
    q = synth.descriptors(3001, 1, True); t = synth.descriptors(20000, 2, True)
    q[:, 2:] = 0; t[:, 2:] = 0        # heavy ties: the cross-tile / cross-rank merge must keep (distance, index) order
    frames = np.stack([synth.triplet(s, 320, 240) for s in (1, 2, 3, 4, 5)])
    out_path = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), q, t, frames, out_path), nprocs=2, join=True)
    g = np.load(out_path)
    ref, k = oracle.knn2_bf(q, t)
    for s in ("0", "1"):
        assert np.array_equal(g["idx" + s], ref["trainIdx"]) and np.array_equal(g["dst" + s], ref["distance"].astype(np.int32)), s
    orbs = [oracle.Orb(300, 1.2, 4, 28, 15) for _ in range(3)]
    for f in range(len(frames)):
        ks, ds = [], []
        for c in range(3):
            n, kp, d = orbs[c].extract(frames[f, c])
            assert g["counts"][f, c] == n
            assert g["kps"][f, c, :n].tobytes() == kp.tobytes() and g["desc"][f, c, :n].tobytes() == d.tobytes()
            ks.append(kp); ds.append(d)
        n, ur, dp, bd, br = oracle.stereo_match(orbs[0], orbs[1], ks[0], ds[0], ks[1], ds[1], 240, 955.40503, 1.0)
        nl = len(ks[0])
        assert g["u_right"][f, :nl].tobytes() == ur.tobytes() and g["depth_left"][f, :nl].tobytes() == dp.tobytes()
