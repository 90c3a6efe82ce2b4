#!/usr/bin/env python
"""configs[3] of BASELINE.json on one GPU: 1280x720 three-camera frames, 5000 ORB features each, extract x3 + L/R stereo,
device-resident (the 4096-frame batch of the config is streamed as steps of --frames frames; shards over GPUs like bench.py).
    python scripts/bench_c4.py [--frames 48] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcvslam_b200.api as A  # noqa: E402
from mcvslam_b200 import synth  # noqa: E402

W, H = 1280, 720


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    rig = A.Rig(bf=955.40503, baseline=1.0, device=0, stream=stream.cuda_stream, nkeypoints=5000, scale_factor=1.2, nlevels=8, ini_th_fast=28, min_th_fast=15)
    cap = rig.cap
    base = [synth.triplet(500 + s, W, H) for s in range(8)]
    frames = np.stack([base[i % 8] for i in range(a.frames)])
    B = a.frames
    d_imgs = torch.from_numpy(frames).to(dev)
    d_kps = torch.empty(B * 3 * cap * 28, dtype=torch.uint8, device=dev); d_desc = torch.empty(B * 3 * cap * 32, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B * 3, dtype=torch.int32, device=dev)
    d_ur = torch.empty(B * cap, dtype=torch.float32, device=dev); d_dp = torch.empty(B * cap, dtype=torch.float32, device=dev)

    def step():
        rig.process_async(d_imgs.data_ptr(), B, W, H, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), d_ur.data_ptr(), d_dp.data_ptr())

    with torch.cuda.stream(stream):
        for _ in range(a.warmup):
            step()
        rig.join(); torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            step()
        rig.join()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    kp = float(d_cnt.float().mean().item())
    print(json.dumps({"workload": "configs[3]: 1280x720 triplets, 5000 ORB x 8 levels x 1.2, extract + L/R stereo", "frames_per_step": B, "steps": a.steps,
                      "ms_per_step": ms / a.steps, "three_camera_frames_per_s": B * a.steps / (ms * 1e-3), "keypoints_per_image": kp,
                      "seconds_for_4096_frames": 4096 / (B * a.steps / (ms * 1e-3))}))


if __name__ == "__main__":
    main()
