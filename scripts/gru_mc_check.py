"""GRU: cluster-multicast weight streaming (WEKWS_GRU_CLUSTER=1) vs the default kernel -- parity and step time."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from wekws_b200 import init_model, model_config, synth
m = synth.randomize_(init_model(model_config('gru'))).eval().cuda()
for B, T in [(512, 1), (100, 5), (37, 3)]:
    x = synth.features(B, T, 80, seed=3).cuda()
    h = torch.randn(2, B, 128, device='cuda')
    outs = {}
    for mc in ("0", "1"):
        os.environ["WEKWS_GRU_CLUSTER"] = mc
        y, c = m(x, h)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200): m(x, h)
        b.record(); torch.cuda.synchronize()
        outs[mc] = (y, c, a.elapsed_time(b) / 200)
    print('B', B, 'T', T, 'maxdiff y %.2e c %.2e' % (float((outs["0"][0] - outs["1"][0]).abs().max()), float((outs["0"][1] - outs["1"][1]).abs().max())),
          'ms default %.4f multicast %.4f' % (outs["0"][2], outs["1"][2]), flush=True)
