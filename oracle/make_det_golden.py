#!/usr/bin/env python
"""Pins the detection-statistics oracle to the REFERENCE's own tool (test infrastructure, not product code).

Generates seeded synthetic posteriors, writes them the way wekws/bin/score.py:128-137 does
('{key} {keyword} {:.6f} {:.6f} ...' per utterance and keyword), writes the matching label file
(json lines with key / txt / duration, compute_det.py:36-50), RUNS the unmodified
/root/reference/wekws/bin/compute_det.py (stdlib only) on them for several window_shift / step
settings, and stores inputs + the stats files it wrote in tests/golden/det_stats.npz.

    python oracle/make_det_golden.py           # needs /root/reference; run in this container only
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TOOL = "/root/reference/wekws/bin/compute_det.py"
KEYWORDS = ["HI_XIAOWEN", "NIHAO_WENWEN"]
SETTINGS = [(50, 0.01), (1, 0.01), (7, 0.05)]          # (window_shift, step)


def synth_posteriors():
    g = torch.Generator().manual_seed(2024)
    B, T, K = 24, 310, 2
    post = torch.sigmoid(torch.randn(B, T, K, generator=g) * 2.5 - 1.0)
    # values whose 6-decimal text lands exactly on thresholds / rounding ties, and saturated scores
    post[0, :6, 0] = torch.tensor([0.5, 0.4999995, 0.5000005, 0.01, 1.0, 0.11])
    post[1, :, 0] = 0.11                                 # max == a multiple of step (ADVICE: float vs double compare)
    post[2] = 0.0
    post[3, :, 1] = 0.299999999
    lens = torch.tensor([310, 309, 300, 200, 120, 51, 50, 49, 2, 1, 310, 310, 77, 310, 150, 310, 33, 310, 310, 250,
                         310, 5, 310, 100], dtype=torch.int32)
    # utterance transcripts: keyword 0, keyword 1 or filler (compute_det.py:44-49)
    kinds = [0, 1, 2, 2, 0, 2, 1, 2, 2, 0, 2, 2, 1, 2, 0, 2, 2, 2, 1, 2, 0, 2, 2, 2]
    durs = [round(0.01 * int(n) + 0.015, 3) for n in lens.tolist()]
    return post, lens, kinds, durs


def write_score_file(path, post, lens):
    with open(path, "w", encoding="utf8") as fout:      # the loop of score.py:128-137
        for i in range(post.shape[0]):
            score = post[i][:int(lens[i])]
            for k, kw in enumerate(KEYWORDS):
                frames = " ".join(["{:.6f}".format(x) for x in score[:, k].tolist()])
                fout.write("{} {} {}\n".format("utt%03d" % i, kw, frames))


def main():
    assert os.path.exists(REF_TOOL), "needs the reference checkout at /root/reference"
    post, lens, kinds, durs = synth_posteriors()
    out = {"post": post.numpy(), "lens": lens.numpy(), "kinds": np.array(kinds, dtype=np.int32),
           "durations": np.array(durs, dtype=np.float64), "settings": np.array(SETTINGS, dtype=np.float64)}
    with tempfile.TemporaryDirectory() as td:
        score = os.path.join(td, "score.txt")
        label = os.path.join(td, "data.list")
        write_score_file(score, post, lens)
        with open(label, "w", encoding="utf8") as f:
            for i, (kd, du) in enumerate(zip(kinds, durs)):
                txt = KEYWORDS[kd].lower() if kd < 2 else "some filler words"      # compute_det upper()s it
                f.write(json.dumps({"key": "utt%03d" % i, "txt": txt, "duration": du}) + "\n")
        for si, (ws, step) in enumerate(SETTINGS):
            for k, kw in enumerate(KEYWORDS):
                stats = os.path.join(td, f"stats_{si}_{k}.txt")
                subprocess.run([sys.executable, REF_TOOL, "--test_data", label, "--keyword", kw, "--score_file", score,
                                "--step", str(step), "--window_shift", str(ws), "--stats_file", stats],
                               check=True, stdout=subprocess.DEVNULL)
                with open(stats, encoding="utf8") as f:
                    out[f"stats_{si}_{k}"] = np.array(f.read())
    dst = os.path.join(ROOT, "tests", "golden", "det_stats.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: (v.shape if hasattr(v, "shape") and v.shape else str(v)[:40]) for k, v in out.items()})


if __name__ == "__main__":
    main()
