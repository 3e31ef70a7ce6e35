#!/usr/bin/env python
"""GRU kernel comparison: FP32 kernel (gru.cu) vs weight-streaming tensor-core kernel (gru_tc.cu) over B, T (CUDA events).
usage: gru_sweep.py [small]   -- `small`: the grid around the dispatch thresholds of wekws_model_forward"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wekws_b200 import init_model, model_config, synth

dev = "cuda:0"
torch.manual_seed(777)
m = synth.randomize_(init_model(model_config("gru"))).eval().to(dev)
out = []
GRID = [(512, 1), (512, 4), (512, 16), (512, 40), (1024, 40), (4096, 40), (8192, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    GRID = [(B, T) for T in (1, 2, 8) for B in (16, 64, 148, 296, 512, 640, 768, 1024, 2048)]
for B, T in GRID:
    x = synth.features(B, T, 80, seed=3).to(dev)
    row = {"B": B, "T": T}
    for prec in ("fp32", "tensor"):
        m.precision = prec
        h = torch.zeros(2, B, 128, device=dev)
        for _ in range(20):
            _, h = m(x, h)
        torch.cuda.synchronize()
        n = 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            _, h = m(x, h)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / n * 1e3
        row[prec + "_us"] = round(us, 2)
        row[prec + "_Mframes_s"] = round(B * T / us, 1)
    out.append(row)
    print(json.dumps(row))
