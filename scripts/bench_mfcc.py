"""Front-end micro-bench: Fbank vs MFCC (fused DCT + lifter epilogue), 10 000 one-second clips of int16 PCM."""
import json, sys, torch
sys.path.insert(0, '/root/repo')
from wekws_b200 import Mfcc, Fbank, synth
pcm = synth.pcm_int16(10000, 16000, seed=1).cuda()
res = {}
for name, fe in (("fbank80", Fbank(80)), ("mfcc80_80", Mfcc(80, 80)), ("mfcc13_23", Mfcc(13, 23))):
    out = fe(pcm); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fe(pcm, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res[name] = {"ms": ms, "frames_per_sec": 10000 * 98 / ms * 1e3, "audio_hours_per_sec": 10000 * 98 / ms * 1e3 / 360000}
print(json.dumps(res))
