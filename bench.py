#!/usr/bin/env python
"""Headline benchmark: three-camera frames/s — per frame 3x ORB extraction (2000 features, 8 levels x 1.2, FAST 28/15) on
640x480 images plus left/right stereo matching (BASELINE.json configs[1]) — on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path over one batch of B synthetic triplets per GPU (weak scaling: every rank owns its own
batch; frames are independent, so there is no data-path collective — only a per-step gather of the per-image keypoint
counts, the job's result directory, over NCCL). One JSON line on rank 0:
  value      device-resident throughput: inputs already in HBM, CUDA-event timed on the engine's stream, max over ranks
  e2e        same metric through the C ABI with HOST buffers (pinned): H2D of the images and D2H of all results inside
             the timed region; e2e.copy_ceiling = the same bytes copied with no kernels at all (what the box allows)
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  batch_sweep  device-resident and synchronous host-in/host-out numbers at 1 / 8 / 64 / 1024 frames per call (batch 1 = the
             reference's own calling pattern, one Frame per call: its latency)
  matching   second headline: Hamming 2-NN descriptor pairs/s, tensor-core path and integer-pipe path, each against its peak
  cpu_baseline  the REFERENCE's own code (oracle/_ref: its translation units compiled unmodified) on the host cores, bounded
             sample, rank 0 at N=1 only: all cores (replica processes), the reference's own threading (one process, its
             ThreadPool(3)), and the oracle port on one thread
--impl reference times oracle/_ref with all host cores as its own arm (falls back to the oracle port when the .so is absent).
--workload c1/c3/c4/c5 measure the other BASELINE configs (one JSON line each, same contract); the driver runs the default.
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
WORKLOAD = "configs[1]: 3-camera rig triplet (left/right/wide) 640x480, 2000 ORB x 8 levels x 1.2, FAST 28/15, extract x3 + L/R stereo match"
# compulsory traffic per three-camera frame = 3 x (input 307200 + pyramid levels 1..7 643332 + 2000 x 60 B
# of keypoints/descriptors) + 16000 B of stereo outputs
ALGO_BYTES_PER_FRAME = 3 * (307200 + 643332 + 2000 * 60) + 16000
PYR_BYTES_PER_IMAGE = 950532          # all 8 levels
# algorithmic bytes per IMAGE of each stage (DESIGN.md "Kernels"): what the stage must read + write at least once
STAGE_BYTES_PER_IMAGE = {
    "pyramid": 307200 + 950532,                 # read input, write 8 levels
    "blur": 2 * 950532,                         # read pyramid, write smoothed pyramid
    "fast_score": 950532 + 4 * 6500,            # read pyramid, write ~6.5k packed scored pixels (round 1 also wrote a dense score map: 2 x 950532)
    "nms_cells": 2 * 4 * 6500,                  # read the scored-pixel lists, write the surviving candidates
    "quadtree": 2 * 4 * 6500 + 4 * 2000,        # read candidates, write them ordered, write 2000 selected
    "orient_desc": 2000 * (31 * 31 + 37 * 37 + 60),  # per keypoint: intensity patch + smoothed patch + 60 B out
    "stereo_match": (2 * 2000 * 60 + 2000 * 16) / 3.0,   # per frame / 3 images
    "stereo_median": 2000 * 12 / 3.0,
}
KERNEL_NAMES = {"pyramid": "k_copy_level0 + 7 x k_resize_march", "blur": "k_gauss7", "fast_score": "k_fast_score",
                "nms_cells": "k_nms_sparse + k_cell_order", "quadtree": "k_octree_prep + k_octree_replay", "orient_desc": "k_orient_desc",
                "stereo_match": "k_stereo_rows + k_stereo_match", "stereo_median": "k_stereo_median"}
I8_DENSE_TOPS = 4500.0    # B200 dense int8 tensor-core peak (NVIDIA's figure; MEASURED_PEAKS.json carries no int8 number)


def int8_measured_equiv_tops():
    """Dense int8 is exactly 2 x bf16 on this chip: twice the MEASURED cuBLAS bf16 burst figure is the measured-equivalent ceiling."""
    pk = peaks()
    return 2.0 * float(pk["bf16_tflops"]) if "bf16_tflops" in pk else None


def make_frames(n_frames, seed0, w=W, h=H):
    from mcvslam_b200 import synth
    base = [synth.triplet(seed0 + s, w, h) for s in range(min(n_frames, 16))]
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


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def time_knn2(A, torch, dev, stream, nq, nt, reps, seed=5, warm=1):
    L = A.lib()
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g).to(dev)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, generator=g).to(dev)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dst = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    run = lambda: A._check(L.mcv_knn2_bf_device(q.data_ptr(), nq, t.data_ptr(), nt, 0, idx.data_ptr(), dst.data_ptr(), stream.cuda_stream))
    for _ in range(warm):
        run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e-3 / reps


def bench_matching(A, torch, dev, stream):
    """Brute-force Hamming 2-NN (mcv_knn2_bf_device) on device-resident random descriptors: configs[0]'s 2000 x 2000, configs[3]'s
    5000 x 5000 and one GPU's query shard of configs[4] (131072 of the 1M queries x all 1M train rows). Two legs: the shipped
    path (tcgen05 int8 GEMM of the +-1 expanded descriptors from 2^23 pairs up, the integer-pipe kernel below; 512 int8 ops per pair against the dense int8
    peak) and the integer-pipe kernel (MCV_KNN_POPC=1; 8 popc32 per pair against the live measured xor+popc peak)."""
    peak, _ = A.popc_peak(8192)
    meq = int8_measured_equiv_tops()
    out = {"unit": "descriptor pairs/s", "int8_measured_equiv_tops": meq,
           "int8_measured_equiv_source": "2 x MEASURED_PEAKS.json bf16_tflops (cuBLAS burst; int8 dense = 2 x bf16 on B200)", "popc32_peak_per_s": peak, "popc_peak_source": "mcv_debug_popc_peak (8 independent xor+popc+add chains per thread, whole GPU)",
           "int8_peak_tops": I8_DENSE_TOPS, "int8_peak_source": "B200 dense int8 tensor-core figure (no measured int8 peak in MEASURED_PEAKS.json)", "cases": []}
    for name, nq, nt, reps in (("configs[0] 2000x2000", 2000, 2000, 50), ("configs[3] pair 5000x5000", 5000, 5000, 50),
                               ("configs[4] shard 131072x1048576", 131072, 1048576, 3)):
        os.environ["MCV_KNN_POPC"] = "0"
        sec = time_knn2(A, torch, dev, stream, nq, nt, reps, warm=3 if nq < 10000 else 1)
        os.environ["MCV_KNN_POPC"] = "1"
        sec_p = time_knn2(A, torch, dev, stream, nq, nt, max(1, reps // 2), warm=3 if nq < 10000 else 1)
        os.environ["MCV_KNN_POPC"] = "0"
        pairs = float(nq) * nt / sec; pairs_p = float(nq) * nt / sec_p
        tensor = nq * nt >= 1 << 23        # dispatch rule of launch_knn2_bf (match_tc_kernels.cu: knn2_tc_usable)
        one_launch = nq >= 256 and nt <= 4096 and not tensor   # match_kernels.cu: wq_usable
        int_kernel = "k_knn2_wq (one launch: warp per query, train set streamed through shared memory)" if one_launch else "k_knn2_bf + k_knn2_merge"
        out["cases"].append({"case": name, "ms": sec * 1e3, "pairs_per_s": pairs, "int8_tops": pairs * 512 / 1e12 if tensor else None,
                             "frac_of_int8_peak": pairs * 512 / 1e12 / I8_DENSE_TOPS if tensor else None,
                             "frac_of_int8_measured_equiv": pairs * 512 / 1e12 / meq if tensor and meq else None,
                             "frac_of_popc_peak": None if tensor else pairs * 8 / peak,
                             "kernel": "k_expand_pm1 + k_knn2_tc (tcgen05.mma kind::i8)" if tensor else int_kernel + " (integer pipe: below 2^23 pairs)",
                             "popc_path": {"ms": sec_p * 1e3, "pairs_per_s": pairs_p, "popc32_per_s": pairs_p * 8, "frac_of_popc_peak": pairs_p * 8 / peak,
                                           "kernel": int_kernel}, "speedup_vs_popc_path": sec_p / sec})
    return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference itself (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def _ref_cfg():
    return (ORB["nkeypoints"], ORB["scale_factor"], ORB["nlevels"], ORB["ini_th_fast"], ORB["min_th_fast"], BF, BASELINE)


def ref_available():
    try:
        from oracle import ref as R
        return R.available()
    except Exception:
        return False


def cpu_port(n_threads, budget_s):
    """The oracle PORT (oracle/orb_oracle.cpp) over a bounded sample with n_threads worker threads."""
    from oracle import oracle as O
    O.build()
    frames = make_frames(8, 100)
    a = (ORB["nkeypoints"], ORB["scale_factor"], ORB["nlevels"], ORB["ini_th_fast"], ORB["min_th_fast"], BF, BASELINE)
    s1, _ = O.bench_frames(frames[:2], *a, n_threads, repeat=max(1, n_threads // 2))
    rate = 2 * max(1, n_threads // 2) / s1
    repeat = max(1, int(rate * budget_s / len(frames)))
    s, _ = O.bench_frames(frames, *a, n_threads, repeat=repeat)
    n = len(frames) * repeat
    return {"value": n / s, "unit": UNIT, "cores": n_threads, "kind": "port",
            "sample": "%d three-camera frames (8 distinct synthetic 640x480 triplets x %d), oracle port, %d threads, %.1f s" % (n, repeat, n_threads, s)}


def cpu_reference(n_procs, budget_s, est_rate_per_proc=11.0):
    """oracle/_ref: n_procs replica processes of the reference's Frame pipeline (each with the reference's own ThreadPool(3))."""
    from oracle import ref as R
    repeat = max(1, int(est_rate_per_proc * budget_s / 8))
    v, n, stages = R.bench_frames(_ref_cfg(), n_procs, 8, 100, W, H, repeat)
    return {"value": v, "unit": UNIT, "cores": 3 * n_procs, "kind": "reference",
            "sample": "%d three-camera frames (8 distinct synthetic 640x480 triplets x %d per replica), %d replica process(es) of the reference's own "
                      "Frame constructor (oracle/_ref: src/Frame.cpp + ORBextractor.cc compiled unmodified), 3 extractor threads each" % (n, repeat, n_procs),
            "reference_stage_s_replica0": stages}


def cpu_baseline(budget_s=12.0):
    """All three variants BASELINE.md §3 names: (a) one thread, (b) the reference's own threading, (c) all host cores."""
    cores = os.cpu_count() or 1
    if not ref_available():
        out = cpu_port(cores, budget_s)
        out["variants"] = {"one_thread_port": cpu_port(1, 4.0)}
        out["note"] = "oracle/_ref/libmcv_ref.so absent: oracle port timed instead"
        return out
    n_procs = max(1, cores // 3)
    out = cpu_reference(n_procs, budget_s)
    out["variants"] = {"one_thread_port": cpu_port(1, 4.0), "reference_threading_one_process": cpu_reference(1, 5.0), "all_cores_port": cpu_port(cores, 5.0)}
    out["host_cores"] = cores
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) with all the host threads it can use. The replica
    processes run once for warmup + steps bounded samples; value = frames of the timed samples / their wall time."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_all = time.time()
    if ref_available():
        n_procs = max(1, cores // 3)
        cpu_reference(n_procs, 2.0)                                                 # warm-up leg (processes start, pages fault in)
        total_budget = max(10.0, min(120.0, 2.0 * args.steps))
        res = cpu_reference(n_procs, total_budget)
    else:
        res = cpu_port(cores, max(10.0, min(120.0, 2.0 * args.steps)))
        res["note"] = "oracle/_ref/libmcv_ref.so absent: oracle port timed instead"
    v = res["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.frames / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": args.frames,
                                            "note": "the reference's CPU implementation on the host cores; a step = a bounded sample of the workload"},
            "cpu_baseline": res,
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


class RigBuffers:
    """Device-resident inputs, N_SETS device output sets (in-flight mcv_rig_process_async calls never share outputs,
    include/mcv_b200.h: the engine rotates calls over at most 6 streams, so a set is reused by the 6th following call at the
    earliest, and that call is ordered behind the set's previous user) and `nf` pinned host output sets for a batch of B frames."""

    N_SETS = 6

    def __init__(self, torch, dev, frames_np, cap, nf):
        self.B = B = len(frames_np)
        self.h_imgs = torch.from_numpy(frames_np).pin_memory()
        self.d_imgs = self.h_imgs.to(dev)
        kb, db = B * 3 * cap * 28, B * 3 * cap * 32
        self.d_out = [dict(kps=torch.empty(kb, dtype=torch.uint8, device=dev), desc=torch.empty(db, dtype=torch.uint8, device=dev),
                           cnt=torch.zeros(B * 3, dtype=torch.int32, device=dev), ur=torch.empty(B * cap, dtype=torch.float32, device=dev),
                           dp=torch.empty(B * cap, dtype=torch.float32, device=dev)) for _ in range(self.N_SETS)]
        self.h_out = [dict(kps=torch.empty(kb, dtype=torch.uint8).pin_memory(), desc=torch.empty(db, dtype=torch.uint8).pin_memory(),
                           cnt=torch.zeros(B * 3, dtype=torch.int32).pin_memory(), ur=torch.empty(B * cap, dtype=torch.float32).pin_memory(),
                           dp=torch.empty(B * cap, dtype=torch.float32).pin_memory()) for _ in range(nf)]
        self.h2d_bytes = int(self.h_imgs.numel())
        self.d2h_bytes = int(kb + db + B * 3 * 4 + 2 * B * cap * 4)
        self.k = 0

    def ptrs(self, o):
        return (o["kps"].data_ptr(), o["desc"].data_ptr(), o["cnt"].data_ptr(), o["ur"].data_ptr(), o["dp"].data_ptr())

    def next_dev(self):
        self.k += 1
        return self.d_out[self.k % self.N_SETS]


def copy_ceiling(torch, dev, buf, steps, barrier):
    """The e2e leg's bytes with NO kernels: per step the pinned image block host->device and a pinned result set device->host,
    on two streams, all ranks at once. Returns seconds per step (wall clock between barriers)."""
    s_in = torch.cuda.Stream(device=dev); s_out = torch.cuda.Stream(device=dev)
    d_in = torch.empty_like(buf.d_imgs)
    o_d, o_h = buf.d_out[0], buf.h_out[0]
    def one():
        with torch.cuda.stream(s_in):
            d_in.copy_(buf.h_imgs, non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in ("kps", "desc", "cnt", "ur", "dp"):
                o_h[k].copy_(o_d[k], non_blocking=True)
    one(); barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    barrier()
    return (time.perf_counter() - t0) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128, help="three-camera frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c1", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matching", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--inflight", type=int, default=4, help="e2e leg: steps kept in flight through mcv_rig_submit (<= 8)")
    ap.add_argument("--chunk", type=int, default=None, help="frames per pipelined chunk inside the engine (default: engine default 32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "c2":
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_workloads
        bench_workloads.main(args, rank, world, local_rank)
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
    NF = max(1, min(8, args.inflight))
    buf = RigBuffers(torch, dev, make_frames(B, 1000 * (rank + 1)), cap, NF)
    gathered = torch.zeros(world * B * 3, dtype=torch.int32, device=dev) if world > 1 else None
    last = {"cnt": buf.d_out[0]["cnt"]}

    def step_device(b=buf):
        o = b.next_dev()
        rig.process_async(b.d_imgs.data_ptr(), b.B, W, H, *b.ptrs(o))
        last["cnt"] = o["cnt"]

    def finish_device(gather=True):
        rig.join()     # order the timing stream after every step's results
        if world > 1 and gather:  # the only collective: gather of the per-image keypoint counts (result directory)
            dist.all_gather_into_tensor(gathered, last["cnt"])

    def step_host(b=buf):
        rig.process_ptrs(b.h_imgs.data_ptr(), b.B, W, H, *b.ptrs(b.h_out[0]), False)

    def run_host_pipelined(n_steps):
        """n_steps through mcv_rig_submit / mcv_rig_wait on pinned HOST buffers, --inflight steps in flight: every step's images are
        copied host->device and every result device->host inside the region; a step counts when its results are on the host
        (its keypoint counts are read)."""
        tickets, total_kp = [], 0
        for k in range(n_steps):
            o = buf.h_out[k % NF]
            if k >= NF:
                rig.wait(tickets[k - NF]); total_kp += int(o["cnt"][0])
            tickets.append(rig.submit(buf.h_imgs.data_ptr(), B, W, H, *buf.ptrs(o)))
        for k in range(max(0, n_steps - NF), n_steps):
            rig.wait(tickets[k]); total_kp += int(buf.h_out[k % NF]["cnt"][0])
        return total_kp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_device(n_steps, step, gather=True):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(n_steps):
            step()
        finish_device(gather)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step_device()
        finish_device()
        barrier()
        ms = timed_device(args.steps, step_device)
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
        # step; (b) the throughput API: mcv_rig_submit / mcv_rig_wait, --inflight (default 4) steps in flight — (b) is the headline e2e
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, args.steps // 2)
        for _ in range(e2e_steps):
            step_host()
        barrier()
        e2e_sync_s = time.perf_counter() - t0
        run_host_pipelined(max(3, 2 * NF + 2))   # every slot the steps rotate over has its buffers before the timed region
        barrier()
        t0 = time.perf_counter()
        run_host_pipelined(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        ceiling_s = copy_ceiling(torch, dev, buf, max(5, args.steps // 3), barrier)
        clocks = sampler.stop() if rank == 0 else None

        # batch sizes 1 / 8 / 64 / 1024: device-resident throughput and the synchronous host-in/host-out call
        sweep = []
        if not args.no_sweep:
            for nb in (1, 8, 64, 1024):
                sb = RigBuffers(torch, dev, make_frames(nb, 1000 * (rank + 1)), cap, 1)
                reps = max(3, min(200, 4096 // nb))
                for _ in range(3):
                    step_device(sb)
                finish_device(False)
                d_ms = timed_device(reps, lambda: step_device(sb), gather=False)
                for _ in range(2):
                    step_host(sb)
                lat = []
                for _ in range(max(3, min(50, 1024 // nb))):
                    t0 = time.perf_counter(); step_host(sb); lat.append(time.perf_counter() - t0)
                sweep.append({"frames_per_call": nb, "device_ms_per_call": d_ms / reps, "sync_call_ms": 1e3 * float(np.median(lat))})
                del sb
            torch.cuda.empty_cache()

        # the reference's SHIPPED configuration (config/extractor.yaml: nkeypoints 2000, nlevels 1, on its 512 x 512 cameras,
        # config/camleft.yaml): not a BASELINE config, reported beside it because it is what the reference itself runs. Fewer FAST
        # candidates than quota on the single level, so the quadtree splits down to single points and the serial heap replay is
        # most of a call.
        shipped = None
        if rank == 0 and not args.no_sweep:
            s_orb = dict(ORB, nlevels=1)
            s_rig = A.Rig(bf=BF, baseline=BASELINE, device=local_rank, stream=stream.cuda_stream, **s_orb)
            shipped = {"config": "config/extractor.yaml as shipped (2000 ORB, nlevels 1, FAST 28/15) on 512x512 triplets", "cases": []}
            for nb in (1, 128):
                sb = RigBuffers(torch, dev, make_frames(nb, 77, 512, 512), s_rig.cap, 1)
                def s_step():
                    s_rig.process_async(sb.d_imgs.data_ptr(), sb.B, 512, 512, *sb.ptrs(sb.next_dev()))
                for _ in range(3):
                    s_step()
                s_rig.join()
                reps = 200 if nb == 1 else 30
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                e0.record(stream)
                for _ in range(reps):
                    s_step()
                s_rig.join()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                d_ms = e0.elapsed_time(e1) / reps
                lat = []
                for _ in range(30 if nb == 1 else 5):
                    t0 = time.perf_counter()
                    s_rig.process_ptrs(sb.h_imgs.data_ptr(), sb.B, 512, 512, *sb.ptrs(sb.h_out[0]), False)
                    lat.append(time.perf_counter() - t0)
                shipped["cases"].append({"frames_per_call": nb, "device_frames_per_s": nb / (d_ms * 1e-3), "device_ms_per_call": d_ms,
                                         "sync_call_ms": 1e3 * float(np.median(lat)), "keypoints_per_image": float(sb.d_out[1]["cnt"].float().mean().item())})
                del sb
            s_rig.close()
            torch.cuda.empty_cache()

    # second headline metric (BASELINE.json): Hamming matches/s = query x train descriptor pairs per second of the brute-force
    # 2-NN, device-resident: tensor-core path against the int8 peak, integer-pipe path against the measured xor+popc peak
    matching = None
    if rank == 0 and not args.no_matching:
        matching = bench_matching(A, torch, dev, stream)

    t = torch.tensor([ms, e2e_s, e2e_sync_s, ceiling_s] + [x for s in sweep for x in (s["device_ms_per_call"], s["sync_call_ms"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, e2e_sync_s, ceiling_s = (float(x) for x in t[:4])
    for i, s in enumerate(sweep):
        s["device_ms_per_call"], s["sync_call_ms"] = float(t[4 + 2 * i]), float(t[5 + 2 * i])
    n_kp = int(last["cnt"].sum().item())

    if rank == 0:
        value = world * B * args.steps / (ms * 1e-3)
        e2e_v = world * B * args.steps / e2e_s
        e2e_sync_v = world * B * e2e_steps / e2e_sync_s
        pk = peaks()
        peak = float(pk.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in pk else "fallback 6650 GB/s (B200_PROFILING.md)"
        dom = max(stage_ms, key=stage_ms.get)
        dom_ms = stage_ms[dom] / max(1, n_calls)
        launches_per_call = {"pyramid": 8, "nms_cells": 2, "quadtree": 2}.get(dom, 1)
        dom_bytes = STAGE_BYTES_PER_IMAGE[dom] * 3 * B
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        for name in ("r04_ncu_traffic.json", "r03_ncu_traffic.json", "r02_ncu_traffic.json", "r01_ncu_traffic.json"):   # DRAM bytes of the same kernel from the committed ncu --set full capture
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["per_image"][dom] * 3 * B
                traffic_src = "profiles/%s (ncu --set full, per launch)" % name
                break
            except Exception:
                pass
        for s in sweep:
            nb = s["frames_per_call"]
            s["device_frames_per_s"] = world * nb / (s["device_ms_per_call"] * 1e-3)
            s["sync_call_frames_per_s"] = world * nb / (s["sync_call_ms"] * 1e-3)
        copy_bytes = buf.h2d_bytes + buf.d2h_bytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "keypoints_per_frame": n_kp / B,
                       "l2": "inputs larger than L2: %d MB of images + %d MB of pyramids per step" % (B * 3 * W * H >> 20, B * 3 * PYR_BYTES_PER_IMAGE >> 20),
                       "collective": "all_gather of per-image keypoint counts per step (N>1 only)",
                       "host_cpus": ("NUMA-local to the GPU: %d CPUs" % len(numa)) if numa else "unbound"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": buf.h2d_bytes, "d2h_bytes_per_step": buf.d2h_bytes,
                    "steps": args.steps, "how": "mcv_rig_submit / mcv_rig_wait on pinned host buffers, %d steps in flight, wall clock; every " % NF +
                                                "step's H2D and D2H inside the region",
                    "sync_call_value": e2e_sync_v, "sync_call_how": "one synchronous mcv_rig_process call per step (%d steps)" % e2e_steps,
                    "copy_ceiling": {"frames_per_s": world * B / ceiling_s, "gbs_aggregate": world * copy_bytes / ceiling_s / 1e9,
                                     "how": "the same pinned H2D + D2H bytes per step with no kernels, all ranks at once (wall clock)"},
                    "frac_of_copy_ceiling": e2e_v / (world * B / ceiling_s)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": KERNEL_NAMES.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "launches_per_step": launches_per_call, "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": dom_ms,
                         "whole_path_frac": (ALGO_BYTES_PER_FRAME * value / world) / 1e9 / peak},
            "stage_ms_per_step": {k: v / max(1, n_calls) for k, v in stage_ms.items()},
            "stage_ms_note": "per-kernel CUDA-event durations from %d extra steps run unchunked on one stream right after the timed region" % n_calls,
            "batch_sweep": sweep,
            "shipped_config": shipped,
            "latency_ms": next((s["sync_call_ms"] for s in sweep if s["frames_per_call"] == 1), None),
            "latency_note": "one synchronous mcv_rig_process call on ONE triplet with host buffers (the reference's Frame-per-call pattern), median",
            "clocks": clocks,
            "matching": matching,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if shipped is not None and ref_available():
                from oracle import ref as R
                cfg1 = (ORB["nkeypoints"], ORB["scale_factor"], 1, ORB["ini_th_fast"], ORB["min_th_fast"], BF, BASELINE)
                v, n, _ = R.bench_frames(cfg1, 1, 8, 77, 512, 512, 4)
                shipped["reference_frames_per_s_one_process"] = v
                shipped["reference_sample"] = "%d triplets, one process of the reference's own Frame constructor (its 3 extractor threads), oracle/_ref" % n
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
