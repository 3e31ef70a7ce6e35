#!/usr/bin/env python
"""GRU kernel comparison: FP32 kernel (gru.cu) vs cluster / tensor-core kernel (gru_tc.cu) over B, T (CUDA events)."""
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
for B, T in [(512, 1), (512, 4), (512, 16), (512, 40), (1024, 40), (4096, 40), (8192, 1)]:
    x = synth.features(B, T, 80, seed=3).to(dev)
    row = {"B": B, "T": T}
    for prec in ("fp32", "auto"):
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
