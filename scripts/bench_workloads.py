"""The other BASELINE.json configs behind `bench.py --workload c1|c3|c4|c5` (same launch and JSON contract as the default):

  c1  configs[0] test/matching_benchmark: two 640x480 images, 2000 ORB each, extract x2 + BF Hamming 2-NN + FilterRatio(0.6) +
      FilterThreshold(46) through the per-image host API (one image pair per step); plus the literal 1 x (1 + N) case
  c3  configs[2] SearchByProjection: 10 k MapPoints projected into 3 cameras, windowed Hamming matching (mcv_project_match)
  c4  configs[3] batch of 4096 synthetic 1280x720 frames, 5000 ORB each: extract + BF 2-NN between consecutive frames, frames
      sharded contiguously over the ranks (one halo frame recomputed per shard boundary, no data-path collective; final gather
      of the match counts)
  c5  configs[4] LargeScaleMatching: all-pairs 2-NN on 2^20 x 2^20 descriptors, queries sharded over the ranks, train set
      broadcast from rank 0 and the result rows all-gathered (both inside the timed region)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup(local_rank, world):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, dev, saved


def _emit(line, saved, dist, world, rank):
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _max_over_ranks(torch, dist, dev, world, vals):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _base(args, world, metric, unit, value, ms_per_step, workload, scaling, extra_cfg=None):
    cfg = {"workload": workload}
    cfg.update(extra_cfg or {})
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg}


# ---------------------------------------------------------------------------------------------------------------------
def run_c5(args, rank, world, local_rank):
    import bench as Bn
    torch, dist, dev, saved = _setup(local_rank, world)
    import mcvslam_b200.api as A
    from mcvslam_b200 import shard
    n = 1 << 20
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    q = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g).to(dev)          # every rank derives the same queries
    t_src = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g).to(dev) if rank == 0 else None
    stream = torch.cuda.current_stream(dev)
    fn = shard.engine_knn2_fn()

    def step():
        t = shard.broadcast_descriptors(t_src, n, dev)                                   # ncclBroadcast of the 32 MiB train set
        return shard.knn2_sharded(q, t, fn, gather=world > 1)                            # local 2-NN + all_gather of the result rows

    steps = max(1, min(args.steps, 10))
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        idx, dst = step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    (ms,) = _max_over_ranks(torch, dist, dev, world, [e0.elapsed_time(e1) / steps])
    # SURVEY 8(d)'s low-entropy variant (every byte in 0..3: 64 random bits per descriptor, distances crowd around 32 and tie in
    # their thousands) — the exact top-2 of the tensor-core epilogue must not depend on the data for its speed
    q2 = torch.randint(0, 4, (n, 32), dtype=torch.uint8, generator=g).to(dev)
    t2_src = torch.randint(0, 4, (n, 32), dtype=torch.uint8, generator=g).to(dev) if rank == 0 else None

    def step_low():
        t2 = shard.broadcast_descriptors(t2_src, n, dev)
        return shard.knn2_sharded(q2, t2, fn, gather=world > 1)

    step_low()
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(2):
        idx2, dst2 = step_low()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    (ms_low,) = _max_over_ranks(torch, dist, dev, world, [e0.elapsed_time(e1) / 2])
    line = None
    if rank == 0:
        pairs = float(n) * n / (ms * 1e-3)
        tops = pairs * 512 / 1e12
        line = _base(args, world, "hamming_pairs_per_s", "descriptor pairs/s", pairs, ms, "configs[4] LargeScaleMatching: all-pairs Hamming 2-NN, "
                     "2^20 x 2^20 256-bit descriptors, queries sharded over the GPUs", "strong",
                     {"collective": "ncclBroadcast of the train set (32 MiB) + all_gather of the result rows, both timed", "queries_per_gpu": n // world})
        line["steps"] = steps
        line["roofline"] = {"bound": "tensor", "kernel": "k_knn2_tc (tcgen05.mma kind::i8)", "achieved": tops / world, "peak": Bn.I8_DENSE_TOPS, "unit": "TOP/s (int8, per GPU)",
                            "frac": tops / world / Bn.I8_DENSE_TOPS, "traffic": None, "peak_source": "B200 dense int8 figure (no measured int8 peak in MEASURED_PEAKS.json)",
                            "algorithmic_ops_per_pair": 512, "peak_measured_equiv": Bn.int8_measured_equiv_tops(),
                            "frac_of_measured_equiv": (tops / world / Bn.int8_measured_equiv_tops()) if Bn.int8_measured_equiv_tops() else None,
                            "peak_measured_equiv_source": "2 x MEASURED_PEAKS.json bf16_tflops (cuBLAS burst): int8 dense = 2 x bf16 on this chip"}
        line["e2e"] = {"value": pairs, "unit": "descriptor pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "descriptors are produced on the device by extraction; the timed region holds the broadcast and the gather"}
        line["gpu_launches"] = 2 * steps
        line["self_check"] = {"nearest_distance_min": int(dst[:, 0].min().item()), "rows": int(idx.shape[0])}
        line["low_entropy_variant"] = {"bytes": "uniform in 0..3 (default_rng-style seed 5 stream continued)", "ms_per_step": ms_low, "pairs_per_s": float(n) * n / (ms_low * 1e-3),
                                       "nearest_distance_min": int(dst2[:, 0].min().item()), "nearest_distance_median": int(dst2[:, 0].median().item()),
                                       "ties_best_equals_second": float((dst2[:, 0] == dst2[:, 1]).float().mean().item())}
    _emit(line, saved, dist, world, rank)


# ---------------------------------------------------------------------------------------------------------------------
def run_c4(args, rank, world, local_rank):
    import bench as Bn
    torch, dist, dev, saved = _setup(local_rank, world)
    import mcvslam_b200.api as A
    from mcvslam_b200 import shard, synth
    W, H, NF, N_FRAMES, CH = 1280, 720, 5000, 4096, 64
    L = A.lib()
    main = torch.cuda.Stream(device=dev)
    pb, pe, fb, fe = shard.consecutive_pairs(N_FRAMES, rank, world)       # this rank's pairs [pb, pe) need frames [fb, fe)
    base = [synth.scene(700 + s, W, H) for s in range(16)]
    # one chunk of CH frames resident at a time; consecutive chunks overlap by one frame (the pair across the boundary). Chunks
    # alternate over TWO extractor handles / streams (like the rig's slots): the latency-bound quadtree of one chunk (level 0 keeps
    # 1086 of ~4600 candidates: a ~0.9 ms serial heap replay) runs beside the stencils of the next.
    h_chunk = torch.from_numpy(np.stack([base[i % 16] for i in range(CH)])).pin_memory()
    pq = torch.arange(0, CH - 1, dtype=torch.int32, device=dev); pt = pq + 1
    n_chunks = max(1, -(-(fe - fb - 1) // (CH - 1)))

    class Lane:
        def __init__(self):
            self.s = torch.cuda.Stream(device=dev)
            self.E = A.ORB(NF, 1.2, 8, 28, 15, device=local_rank, stream=self.s.cuda_stream)
            self.cap = self.E.max_keypoints(0, W, H)
            cap = self.cap
            self.d_chunk = h_chunk.to(dev)
            self.d_kps = torch.empty(CH * cap * 28, dtype=torch.uint8, device=dev); self.d_desc = torch.empty((CH, cap, 32), dtype=torch.uint8, device=dev)
            self.d_cnt = torch.zeros(CH, dtype=torch.int32, device=dev)
            self.d_idx = torch.empty((CH - 1, cap, 2), dtype=torch.int32, device=dev); self.d_dst = torch.empty((CH - 1, cap, 2), dtype=torch.int32, device=dev)
            self.h_idx = torch.empty((CH - 1, cap, 2), dtype=torch.int32).pin_memory(); self.h_dst = torch.empty((CH - 1, cap, 2), dtype=torch.int32).pin_memory()
            self.h_cnt = torch.zeros(CH, dtype=torch.int32).pin_memory()

        def chunk(self, host):
            with torch.cuda.stream(self.s):
                if host:
                    self.d_chunk.copy_(h_chunk, non_blocking=True)
                A._check(L.mcv_orb_extract_batch_async(self.E._h, self.d_chunk.data_ptr(), CH, W, H, self.d_kps.data_ptr(), self.d_desc.data_ptr(),
                                                       self.d_cnt.data_ptr(), self.cap))
                A._check(L.mcv_knn2_pairs_device(self.d_desc.data_ptr(), self.d_cnt.data_ptr(), CH, self.cap, pq.data_ptr(), pt.data_ptr(), CH - 1,
                                                 self.d_idx.data_ptr(), self.d_dst.data_ptr(), self.s.cuda_stream))
                if host:
                    self.h_idx.copy_(self.d_idx, non_blocking=True); self.h_dst.copy_(self.d_dst, non_blocking=True); self.h_cnt.copy_(self.d_cnt, non_blocking=True)

    lanes = [Lane(), Lane()]
    cap = lanes[0].cap
    d_cnt, d_desc, d_idx, d_dst, h_idx = lanes[0].d_cnt, lanes[0].d_desc, lanes[0].d_idx, lanes[0].d_dst, lanes[0].h_idx

    def job(host):
        for ln in lanes:
            ln.s.wait_stream(main)
        for k in range(n_chunks):
            lanes[k % 2].chunk(host)
        for ln in lanes:
            main.wait_stream(ln.s)

    with torch.cuda.stream(main):
        for ln in lanes:
            ln.chunk(False); ln.chunk(False); ln.chunk(False)
        job(False)                                                           # one untimed pass over the whole batch (clocks, pools, L2)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        PASSES = 3
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(PASSES):
            job(False)
            tot = d_cnt.sum().reshape(1)
            if world > 1:
                dist.all_reduce(tot)                                         # final gather of the result directory
        e1.record(main)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / PASSES
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter(); job(True); torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e2e_s = time.perf_counter() - t0
        # share of the step spent matching (one lane alone)
        e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
        e2.record(main)
        for _ in range(5):
            A._check(L.mcv_knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), CH, cap, pq.data_ptr(), pt.data_ptr(), CH - 1, d_idx.data_ptr(), d_dst.data_ptr(),
                                             main.cuda_stream))
        e3.record(main); torch.cuda.synchronize(dev)
        match_ms = e2.elapsed_time(e3) / 5
    ms, e2e_s = _max_over_ranks(torch, dist, dev, world, [ms, e2e_s])
    line = None
    if rank == 0:
        kp = float(d_cnt.float().mean().item())
        v = N_FRAMES / (ms * 1e-3)
        pairs_per_chunk = float(((d_cnt[:-1].double()) * (d_cnt[1:].double())).sum().item())
        line = _base(args, world, "frames_per_s", "frames/s", v, ms, "configs[3]: batch of 4096 synthetic 1280x720 frames, 5000 ORB x 8 levels x 1.2 each, extract + BF "
                     "Hamming 2-NN between consecutive frames, frames sharded over the GPUs", "strong",
                     {"frames_per_chunk": CH, "chunks_per_gpu": n_chunks, "keypoints_per_frame": kp, "halo": "one frame recomputed per chunk / shard boundary",
                      "streams": "chunks alternate over two extractor handles / streams",
                      "collective": "all_reduce of the keypoint count (result directory) at the end"})
        line["steps"] = 3
        line["warmup"] = 1
        line["seconds_for_4096_frames"] = ms * 1e-3
        line["e2e"] = {"value": N_FRAMES / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(h_chunk.numel()) * n_chunks,
                       "d2h_bytes_per_step": int(h_idx.numel() * 8 + CH * 4) * n_chunks, "how": "pinned images in, match tables (idx, dist) and counts out, per chunk"}
        line["matching_share"] = {"ms_per_chunk": match_ms, "pairs_per_s": pairs_per_chunk / (match_ms * 1e-3), "chunk_ms": ms / n_chunks}
        hbm = float(Bn.peaks().get("hbm_gbs", 6650.0))
        algo = 921600 + 1931488 + 5000 * 60                                  # compulsory bytes per 1280x720 image
        line["roofline"] = {"bound": "hbm", "kernel": "whole extraction path", "achieved": algo * v / world / 1e9, "peak": hbm, "unit": "GB/s", "frac": algo * v / world / 1e9 / hbm,
                            "traffic": None, "algorithmic_bytes_per_frame": algo}
        line["gpu_launches"] = int(n_chunks * 30)
    _emit(line, saved, dist, world, rank)


# ---------------------------------------------------------------------------------------------------------------------
def run_c3(args, rank, world, local_rank):
    torch, dist, dev, saved = _setup(local_rank, world)
    import mcvslam_b200.api as A
    from mcvslam_b200 import synth
    W, H, N_MP = 640, 480, 10000
    rng = np.random.default_rng(6 + rank)
    trip = synth.triplet(55 + rank)
    E = A.ORB(2000, 1.2, 8, 28, 15)
    cams = []
    fx = fy = np.float32(955.40503 * 640 / 512); cx, cy = np.float32(320), np.float32(240)
    R = np.eye(3, dtype=np.float32); tt = np.array([0.02, -0.01, 0.03], np.float32)
    for c in range(3):
        n, k, d = E.Extract(trip[c])
        src = rng.integers(0, n, N_MP)
        z = rng.uniform(2, 50, N_MP).astype(np.float32)
        u = k["x"][src] + rng.normal(0, 2.0, N_MP).astype(np.float32); v = k["y"][src] + rng.normal(0, 2.0, N_MP).astype(np.float32)
        pw = (np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32) - tt).astype(np.float32)
        md = d[src].copy()
        for i in range(0, N_MP, 3):                       # a third of the MapPoint descriptors carry random bit flips
            b = rng.choice(256, int(rng.integers(0, 40)), replace=False)
            np.bitwise_xor.at(md[i], b // 8, (1 << (b % 8)).astype(np.uint8))
        cams.append((k, d, pw, md, k["octave"][src].astype(np.int32)))

    def step(r_th=5.0):
        tot = 0
        for k, d, pw, md, lvl in cams:
            tot += A.ProjectBunchMapPoints(k, d, W, H, E.mvScaleFactor, R, tt, [fx, fy, cx, cy], pw, md, lvl, r_th)[0]
        return tot

    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        dist.barrier()
    steps = max(5, min(args.steps, 200))
    t0 = time.perf_counter()
    for _ in range(steps):
        matched = step()
    s = time.perf_counter() - t0
    (s,) = _max_over_ranks(torch, dist, dev, world, [s])
    line = None
    if rank == 0:
        v = world * 3 * N_MP * steps / s
        line = _base(args, world, "projected_mappoints_per_s", "MapPoints/s", v, 1e3 * s / steps, "configs[2] Tracker SearchByProjection: 10 k MapPoints projected into each "
                     "of 3 cameras (2000 ORB keypoints, 640x480), 30x30 grid window + Hamming 2-NN + ratio 0.6 + threshold 46", "weak",
                     {"matched_per_step": int(matched), "how": "three synchronous mcv_project_match calls per step, HOST buffers in and out (replicas only: one rig per GPU)"})
        line["steps"] = steps
        line["e2e"] = {"value": v, "unit": "MapPoints/s", "h2d_bytes_per_step": int(sum(k.nbytes + d.nbytes + pw.nbytes + md.nbytes + lvl.nbytes for k, d, pw, md, lvl in cams)),
                       "d2h_bytes_per_step": 3 * N_MP * 8}
        line["gpu_launches"] = 3 * steps
        try:
            from oracle import ref as Rf
            if Rf.available():
                ro = Rf.Orb()
                k, d, pw, md, lvl = cams[0]
                obj = Rf.Obj(ro, k, d, W, H, [fx, fy, cx, cy], R, tt)
                t1 = time.perf_counter(); obj.project_match(pw, md, lvl, 5.0); cs = time.perf_counter() - t1
                line["cpu_baseline"] = {"value": N_MP / cs, "unit": "MapPoints/s", "cores": 1, "kind": "reference",
                                        "sample": "10 k MapPoints through the reference's Object::ProjectBunchMapPoints (oracle/_ref), one camera, one thread"}
        except Exception as e:          # the baseline is informative; never fail the GPU line on it
            line["cpu_baseline"] = {"error": str(e)}
    _emit(line, saved, dist, world, rank)


# ---------------------------------------------------------------------------------------------------------------------
def run_c1(args, rank, world, local_rank):
    torch, dist, dev, saved = _setup(local_rank, world)
    import mcvslam_b200.api as A
    from mcvslam_b200 import synth
    a, b = synth.scene(1000 + 2 * rank), synth.scene(1001 + 2 * rank)
    E = A.ORB(2000, 1.2, 8, 28, 15)

    def step():
        n1, k1, d1 = E.Extract(a); n2, k2, d2 = E.Extract(b)
        m = A.Matcher.KnnMatch(d1, d2).FilterRatio(0.6).FilterThreshold(46)
        return n1, n2, len(m), d1, d2

    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        dist.barrier()
    steps = max(5, min(args.steps, 200))
    t0 = time.perf_counter()
    for _ in range(steps):
        n1, n2, nm, d1, d2 = step()
    s = time.perf_counter() - t0
    lit = []
    for _ in range(50):                                   # the literal reference case: one query row against the whole other image
        t1 = time.perf_counter(); A.Matcher.KnnMatch(d1[:1], d2); lit.append(time.perf_counter() - t1)
    (s,) = _max_over_ranks(torch, dist, dev, world, [s])
    line = None
    if rank == 0:
        v = world * steps / s
        line = _base(args, world, "matching_benchmark_image_pairs_per_s", "image pairs/s", v, 1e3 * s / steps, "configs[0] test/matching_benchmark: two synthetic 640x480 images, "
                     "2000 ORB x 8 levels x 1.2 each, extract x2 + BF Hamming 2-NN + FilterRatio(0.6) + FilterThreshold(46)", "weak",
                     {"keypoints": [int(n1), int(n2)], "matches_kept": int(nm), "descriptor_pairs_per_step": int(n1) * int(n2),
                      "how": "per-image host API (mcv_orb_extract x2, mcv_knn2_bf, host filters), synchronous calls, HOST buffers"})
        line["steps"] = steps
        line["e2e"] = {"value": v, "unit": "image pairs/s", "h2d_bytes_per_step": 2 * 640 * 480 + (int(n1) + int(n2)) * 32, "d2h_bytes_per_step": (int(n1) + int(n2)) * 60 + int(n1) * 32}
        line["literal_1_x_N_knn_ms"] = 1e3 * float(np.median(lit))
        line["gpu_launches"] = steps * 50
        try:
            from oracle import ref as Rf
            if Rf.available():
                ro = Rf.Orb()
                t1 = time.perf_counter()
                for _ in range(5):
                    _, _, r1 = ro.extract(a); _, _, r2 = ro.extract(b)
                    Rf.filter_threshold(Rf.filter_ratio(Rf.knn2_bf(r1, r2)[0], 0.6), 46)
                cs = (time.perf_counter() - t1) / 5
                line["cpu_baseline"] = {"value": 1.0 / cs, "unit": "image pairs/s", "cores": 1, "kind": "reference",
                                        "sample": "5 image pairs through the reference's ORB::Extract + Matcher::KnnMatch + filters (oracle/_ref), one thread"}
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}
    _emit(line, saved, dist, world, rank)


def main(args, rank, world, local_rank):
    {"c1": run_c1, "c3": run_c3, "c4": run_c4, "c5": run_c5}[args.workload](args, rank, world, local_rank)
