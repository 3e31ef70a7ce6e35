#!/usr/bin/env python
"""Micro-benchmark of the fused Fbank(+CMVN) kernel: B clips x N samples of int16 PCM -> log-mel.
Prints one JSON line with frames/s and the achieved fraction of the HBM roofline
(algorithmic bytes: 160 new samples x 2 B in + 80 x 4 B out = 640 B/frame)."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wekws_b200 import Fbank, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=10000)
    ap.add_argument("--samples", type=int, default=16000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--f32", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    fb = Fbank(80)
    base = synth.pcm_int16(64, a.samples, seed=1234)
    sets = []
    for s in range(3):      # 3 x (320 MB in + 313 MB out) > L2
        p = base.repeat((a.clips + 63) // 64, 1)[:a.clips].roll(s, 0).contiguous()
        sets.append((p.float() if a.f32 else p).to(dev))
    m = fb.num_frames(a.samples)
    out = torch.empty(a.clips, m, 80, device=dev)
    mean = torch.zeros(80, device=dev)
    istd = torch.ones(80, device=dev)
    for i in range(3):
        fb(sets[i % 3], mean=mean, istd=istd, out=out)
    torch.cuda.synchronize()
    ts = []
    for i in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fb(sets[i % 3], mean=mean, istd=istd, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.mean(ts)
    frames = a.clips * m
    in_b = 4 if a.f32 else 2
    bpf = 160 * in_b + 80 * 4
    peak = 6566.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    gbs = frames * bpf / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "fbank_kernel", "clips": a.clips, "samples": a.samples, "frames": frames,
                      "ms": ms, "frames_per_sec": frames / (ms * 1e-3), "audio_hours_per_sec": frames / (ms * 1e-3) / 360000.0,
                      "pcm": "f32" if a.f32 else "s16", "algorithmic_bytes_per_frame": bpf, "achieved_gbs": gbs,
                      "peak_gbs": peak, "frac": gbs / peak}))


if __name__ == "__main__":
    main()
