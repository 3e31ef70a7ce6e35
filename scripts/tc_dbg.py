"""Tensor-core path vs FP32-FMA path of the same model on random inputs: python scripts/tc_dbg.py <model> [B,T ...]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from wekws_b200 import init_model, model_config, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'ds_tcn'
shapes = [tuple(int(v) for v in s.split(',')) for s in sys.argv[2:]] or [(2, 40), (5, 40), (300, 40), (1024, 40), (3, 100), (7, 8), (200, 17), (2, 300)]
m = synth.randomize_(init_model(model_config(name))).eval().cuda()
for B, T in shapes:
    x = synth.features(B, T, 80).cuda()
    outs = {}
    for prec in ("fp32", "auto"):
        m.precision = prec
        y1, c1 = m(x)
        y2, c2 = m(x, c1)
        y3, c3 = m(x, c2.clone())
        torch.cuda.synchronize()
        outs[prec] = (y1, c1, y2, c2, y3)
    d = [float((a - b).abs().max()) for a, b in zip(outs["fp32"], outs["auto"])]
    print(name, 'B', B, 'T', T, 'tc', m.uses_tensor_cores(T), 'maxdiff y1 c1 y2 c2 y3', ['%.2e' % v for v in d], flush=True)
