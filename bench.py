#!/usr/bin/env python
"""Headline benchmark: three-camera frames/s — per frame 3x ORB extraction (2000 features, 8 levels x 1.2, FAST 28/15) on
640x480 images plus left/right stereo matching (BASELINE.json configs[1]) — on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path over one batch of B synthetic triplets per GPU (weak scaling: every rank owns its own
batch; frames are independent, so there is no data-path collective — only a per-step gather of the per-image keypoint
counts, the job's result directory, over NCCL). One JSON line on rank 0:
  value      device-resident throughput: inputs already in HBM, CUDA-event timed on the engine's stream, max over ranks
  e2e        same metric through the C ABI with HOST buffers (pinned): H2D of the images and D2H of all results inside
             the timed region
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (port of the reference's algorithm) on the host cores, bounded sample, rank 0 at N=1 only
--impl reference times the CPU oracle on this box's host cores with all threads (the reference itself cannot be built
here: no OpenCV C++/Eigen/Boost/ROS/pyp); it is the only place besides cpu_baseline where bench.py executes oracle/.
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

W, H = 640, 480
ORB = dict(nkeypoints=2000, scale_factor=1.2, nlevels=8, ini_th_fast=28, min_th_fast=15)
BF, BASELINE = 955.40503, 1.0
METRIC = "three_camera_frames_per_s"
UNIT = "frames/s"
# compulsory traffic per three-camera frame = 3 x (input 307200 + pyramid levels 1..7 643332 + 2000 x 60 B
# of keypoints/descriptors) + 16000 B of stereo outputs
ALGO_BYTES_PER_FRAME = 3 * (307200 + 643332 + 2000 * 60) + 16000
PYR_BYTES_PER_IMAGE = 950532          # all 8 levels
# algorithmic bytes per IMAGE of each stage (DESIGN.md "Kernels"): what the stage must read + write at least once
STAGE_BYTES_PER_IMAGE = {
    "pyramid": 307200 + 950532,                 # read input, write 8 levels
    "blur": 2 * 950532,                         # read pyramid, write smoothed pyramid
    "fast_score": 2 * 950532,                   # read pyramid, write the score map
    "nms_cells": 950532 + 4 * 6500,             # read score map, write ~6.5k packed candidates
    "quadtree": 2 * 4 * 6500 + 4 * 2000,        # read candidates, write them ordered, write 2000 selected
    "orient_desc": 2000 * (31 * 31 + 37 * 37 + 60),  # per keypoint: intensity patch + smoothed patch + 60 B out
    "stereo_match": (2 * 2000 * 60 + 2000 * 16) / 3.0,   # per frame / 3 images
    "stereo_median": 2000 * 12 / 3.0,
}


def make_frames(n_frames, seed0):
    from mcvslam_b200 import synth
    base = [synth.triplet(seed0 + s, W, H) for s in range(min(n_frames, 16))]
    return np.stack([base[i % len(base)] for i in range(n_frames)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bench_matching(A, torch, dev, stream):
    """Brute-force Hamming 2-NN (mcv_knn2_bf_device) on device-resident random descriptors: configs[0]'s 2000 x 2000 and one
    GPU's query shard of configs[4] (131072 of the 1M queries x all 1M train rows). pairs/s, and popc32/s against the live
    measured xor+popc peak (8 popc32 per 256-bit pair)."""
    L = A.lib()
    peak, _ = A.popc_peak(8192)
    out = {"unit": "descriptor pairs/s", "popc32_peak_per_s": peak, "peak_source": "mcv_debug_popc_peak (8 independent xor+popc+add chains per thread, whole GPU)", "cases": []}
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    for name, nq, nt, reps in (("configs[0] 2000x2000", 2000, 2000, 50), ("configs[4] shard 131072x1048576", 131072, 1048576, 2)):
        q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g).to(dev)
        t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, generator=g).to(dev)
        idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dst = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        run = lambda: A._check(L.mcv_knn2_bf_device(q.data_ptr(), nq, t.data_ptr(), nt, 0, idx.data_ptr(), dst.data_ptr(), stream.cuda_stream))
        for _ in range(3 if nq < 10000 else 1):
            run()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        pairs = float(nq) * nt / sec
        out["cases"].append({"case": name, "ms": sec * 1e3, "pairs_per_s": pairs, "popc32_per_s": pairs * 8, "frac_of_popc_peak": pairs * 8 / peak,
                             "self_match_check": None})
    return out


def cpu_baseline(n_threads, budget_s=15.0):
    """The CPU oracle (a port of the reference's algorithm, oracle/) on the host cores over a bounded sample."""
    from oracle import oracle as O
    O.build()
    frames = make_frames(8, 100)
    s1, _ = O.bench_frames(frames[:2], ORB["nkeypoints"], ORB["scale_factor"], ORB["nlevels"], ORB["ini_th_fast"], ORB["min_th_fast"],
                           BF, BASELINE, n_threads, repeat=max(1, n_threads // 2))
    done = 2 * max(1, n_threads // 2)
    rate = done / s1
    repeat = max(1, int(rate * budget_s / len(frames)))
    s, _ = O.bench_frames(frames, ORB["nkeypoints"], ORB["scale_factor"], ORB["nlevels"], ORB["ini_th_fast"], ORB["min_th_fast"],
                          BF, BASELINE, n_threads, repeat=repeat)
    n = len(frames) * repeat
    return {"value": n / s, "unit": UNIT, "cores": n_threads, "kind": "port",
            "sample": "%d three-camera frames (8 distinct synthetic 640x480 triplets x %d), %d threads, %.1f s" % (n, repeat, n_threads, s)}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    t_all = time.time()
    for _ in range(args.warmup):
        cpu_baseline(cores, budget_s=1.0)
    per_step = max(1.0, min(20.0, 150.0 / max(1, args.steps)))
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(cores, budget_s=per_step)
        vals.append(last["value"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.frames / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": {"workload": "configs[1]: 3-camera rig triplet 640x480, 2000 ORB x 8 levels x 1.2, extract + L/R stereo",
                                            "frames_per_step": args.frames, "note": "CPU oracle (port of the reference algorithm); step = bounded sample"},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t_all}
    print(json.dumps(line), flush=True)


def pin_to_gpu_numa(local_rank):
    """Binds this rank to the CPUs NVML reports as local to its GPU, BEFORE the pinned host buffers are allocated (first touch
    puts them on that NUMA node), so that the host-in/host-out leg of N ranks does not cross sockets. Returns the CPU list or
    None when NVML has nothing to say (or the cgroup leaves no such CPU)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in allowed if (words[c // 64] >> (c % 64)) & 1)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128, help="three-camera frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matching", action="store_true")
    ap.add_argument("--inflight", type=int, default=3, help="e2e leg: steps kept in flight through mcv_rig_submit (<= 8)")
    ap.add_argument("--chunk", type=int, default=None, help="frames per pipelined chunk inside the engine (default: engine default 32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import mcvslam_b200.api as A

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa(local_rank) if os.environ.get("MCV_BENCH_NUMA", "1") != "0" else None
    dev = torch.device("cuda", local_rank)
    # NCCL prints its version banner on stdout at the first collective; rank 0's stdout must carry ONE JSON line, so fd 1
    # points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.frames
    stream = torch.cuda.Stream(device=dev)
    rig = A.Rig(bf=BF, baseline=BASELINE, device=local_rank, stream=stream.cuda_stream, **ORB)
    cap = rig.cap
    if args.chunk is not None:
        rig.set_chunk_frames(args.chunk)
    frames_np = make_frames(B, 1000 * (rank + 1))
    # pinned host buffers (e2e leg) and device-resident copies (kernel leg)
    h_imgs = torch.from_numpy(frames_np).pin_memory()
    d_imgs = h_imgs.to(dev)
    kp_bytes = B * 3 * cap * 28
    d_kps = torch.empty(kp_bytes, dtype=torch.uint8, device=dev); d_desc = torch.empty(B * 3 * cap * 32, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B * 3, dtype=torch.int32, device=dev)
    d_ur = torch.empty(B * cap, dtype=torch.float32, device=dev); d_dp = torch.empty(B * cap, dtype=torch.float32, device=dev)
    # --inflight sets of pinned result buffers: the e2e leg keeps that many steps in flight (step k's results are read while the next ones run)
    h_out = []
    NF = max(1, min(8, args.inflight))
    for _ in range(NF):
        h_out.append(dict(kps=torch.empty(kp_bytes, dtype=torch.uint8).pin_memory(), desc=torch.empty(B * 3 * cap * 32, dtype=torch.uint8).pin_memory(),
                          cnt=torch.zeros(B * 3, dtype=torch.int32).pin_memory(), ur=torch.empty(B * cap, dtype=torch.float32).pin_memory(),
                          dp=torch.empty(B * cap, dtype=torch.float32).pin_memory()))
    h_kps, h_desc, h_cnt, h_ur, h_dp = (h_out[0][k] for k in ("kps", "desc", "cnt", "ur", "dp"))
    gathered = torch.zeros(world * B * 3, dtype=torch.int32, device=dev) if world > 1 else None

    def step_device():
        rig.process_async(d_imgs.data_ptr(), B, W, H, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), d_ur.data_ptr(), d_dp.data_ptr())

    def finish_device():
        rig.join()     # order the timing stream after every step's results
        if world > 1:  # the only collective: gather of the per-image keypoint counts (result directory)
            dist.all_gather_into_tensor(gathered, d_cnt)

    def step_host():
        rig.process_ptrs(h_imgs.data_ptr(), B, W, H, h_kps.data_ptr(), h_desc.data_ptr(), h_cnt.data_ptr(), h_ur.data_ptr(), h_dp.data_ptr(), False)

    def run_host_pipelined(n_steps):
        """n_steps through mcv_rig_submit / mcv_rig_wait on pinned HOST buffers, --inflight steps in flight: every step's images are
        copied host->device and every result device->host inside the region; a step counts when its results are on the host
        (its keypoint counts are read)."""
        tickets, total_kp = [], 0
        for k in range(n_steps):
            o = h_out[k % NF]
            if k >= NF:
                rig.wait(tickets[k - NF]); total_kp += int(o["cnt"][0])
            tickets.append(rig.submit(h_imgs.data_ptr(), B, W, H, o["kps"].data_ptr(), o["desc"].data_ptr(), o["cnt"].data_ptr(),
                                      o["ur"].data_ptr(), o["dp"].data_ptr()))
        for k in range(max(0, n_steps - NF), n_steps):
            rig.wait(tickets[k]); total_kp += int(h_out[k % NF]["cnt"][0])
        return total_kp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step_device()
        finish_device()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        finish_device()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = rig.last_launches() * args.steps
        # per-kernel durations: same steps again with the engine's stage events on (whole batch on one stream, no chunk overlap)
        rig.set_profiling(True)
        for _ in range(max(3, min(10, args.steps))):
            step_device()
        finish_device()
        barrier()
        stage_ms, n_calls = rig.stage_ms()
        rig.set_profiling(False)
        # e2e leg: host buffers through the C ABI, copies inside the timed region. (a) one synchronous mcv_rig_process call per
        # step; (b) the throughput API: mcv_rig_submit / mcv_rig_wait, --inflight (default 3) steps in flight — (b) is the headline e2e
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, args.steps // 2)
        for _ in range(e2e_steps):
            step_host()
        barrier()
        e2e_sync_s = time.perf_counter() - t0
        run_host_pipelined(3)
        barrier()
        t0 = time.perf_counter()
        run_host_pipelined(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 else None

    # second headline metric (BASELINE.json): Hamming matches/s = query x train descriptor pairs per second of the brute-force
    # 2-NN kernel, device-resident, against the measured xor+popc peak of this GPU (integer-pipe roofline)
    matching = None
    if rank == 0 and not args.no_matching:
        matching = bench_matching(A, torch, dev, stream)

    t = torch.tensor([ms, e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])
    n_kp = int(d_cnt.sum().item())

    if rank == 0:
        value = world * B * args.steps / (ms * 1e-3)
        e2e_v = world * B * args.steps / e2e_s
        e2e_sync_v = world * B * e2e_steps / e2e_sync_s
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dom = max(stage_ms, key=stage_ms.get)
        dom_ms = stage_ms[dom] / max(1, n_calls)
        launches_per_call = {"pyramid": 8, "nms_cells": 2, "quadtree": 2}.get(dom, 1)
        dom_bytes = STAGE_BYTES_PER_IMAGE[dom] * 3 * B
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of the same kernel from the committed ncu --set full capture, scaled to this launch's images
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["per_image"][dom] * 3 * B
        except Exception:
            pass
        kernel_names = {"pyramid": "k_copy_level0 + 7 x k_resize_march", "blur": "k_gauss7", "fast_score": "k_fast_score",
                        "nms_cells": "k_nms_sparse + k_cell_order", "quadtree": "k_octree_prep + k_octree_replay", "orient_desc": "k_orient_desc",
                        "stereo_match": "k_stereo_rows + k_stereo_match", "stereo_median": "k_stereo_median"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "configs[1]: 3-camera rig triplet (left/right/wide) 640x480, 2000 ORB x 8 levels x 1.2, FAST 28/15, "
                                   "extract + L/R stereo match", "frames_per_step_per_gpu": B, "keypoints_per_frame": n_kp / B,
                       "l2": "inputs larger than L2: %d MB of images + %d MB of pyramids per step" % (B * 3 * W * H >> 20, B * 3 * PYR_BYTES_PER_IMAGE >> 20),
                       "collective": "all_gather of per-image keypoint counts per step (N>1 only)",
                       "host_cpus": ("NUMA-local to the GPU: %d CPUs" % len(numa)) if numa else "unbound"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h_imgs.numel()),
                    "d2h_bytes_per_step": int(h_kps.numel() + h_desc.numel() + h_cnt.numel() * 4 + h_ur.numel() * 4 + h_dp.numel() * 4),
                    "steps": args.steps, "how": "mcv_rig_submit / mcv_rig_wait on pinned host buffers, %d steps in flight, wall clock; every " % NF +
                                                "step's H2D and D2H inside the region",
                    "sync_call_value": e2e_sync_v, "sync_call_how": "one synchronous mcv_rig_process call per step (%d steps)" % e2e_steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kernel_names.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/r01_ncu_traffic.json (ncu --set full, per launch)", "peak_source": peak_src, "launches_per_step": launches_per_call,
                         "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": dom_ms,
                         "whole_path_frac": (ALGO_BYTES_PER_FRAME * value / world) / 1e9 / peak},
            "stage_ms_per_step": {k: v / max(1, n_calls) for k, v in stage_ms.items()},
            "stage_ms_note": "per-kernel CUDA-event durations from %d extra steps run unchunked on one stream right after the timed region" % n_calls,
            "clocks": clocks,
            "matching": matching,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
