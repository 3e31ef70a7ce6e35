import sys, torch
sys.path.insert(0, '/root/repo')
from wekws_b200 import init_model, model_config, synth
m = synth.randomize_(init_model(model_config('tcn'))).eval().cuda()
for B, T in [(2, 40), (5, 40), (300, 40), (1024, 40), (3, 100), (7, 8), (200, 17)]:
    x = synth.features(B, T, 80).cuda()
    outs = {}
    for prec in ("fp32", "auto"):
        m.precision = prec
        y1, c1 = m(x)
        y2, c2 = m(x, c1)
        torch.cuda.synchronize()
        outs[prec] = (y1, c1, y2, c2)
    d = [float((a - b).abs().max()) for a, b in zip(outs["fp32"], outs["auto"])]
    print('B', B, 'T', T, 'tc', m.uses_tensor_cores(T), 'maxdiff y1 c1 y2 c2', d, flush=True)
