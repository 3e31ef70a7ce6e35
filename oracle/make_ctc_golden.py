#!/usr/bin/env python
"""Golden vectors for the CTC prefix beam search (test infrastructure): runs the REFERENCE's own
wekws/model/loss.py ctc_prefix_beam_search (importable in this container) on seeded synthetic posteriors that spell
keyword token sequences with repeats, blanks and competing tokens, and stores inputs + the returned hypotheses in
tests/golden/ctc.npz.      python oracle/make_ctc_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from wekws.model.loss import ctc_prefix_beam_search  # noqa: E402

V = 48
KEYWORDS = {"hi_xiaowen": [5, 9, 17, 23], "nihao_wenwen": [31, 7, 23, 23]}       # second one has a repeated token
TOKENSET = {0, 5, 7, 9, 17, 23, 31}


def synth_probs(T, seed, sharp):
    """Peaky CTC-like softmax posteriors: every frame has one dominant token -- blank unless a (keyword or filler)
    token sequence is being spelled, with repeats, blank gaps and an occasional strong competitor.  The reference's
    search multiplies raw probabilities (no logs), so non-peaky input drives every path below its 1e-6 isclose gate."""
    g = torch.Generator().manual_seed(seed)
    seqs = [KEYWORDS["hi_xiaowen"], KEYWORDS["nihao_wenwen"], [5, 9, 40, 17], [23, 23, 7]]
    logits = torch.randn(T, V, generator=g) * 0.3
    dom = torch.zeros(T, dtype=torch.long)
    t = int(torch.randint(0, 6, (1,), generator=g))
    while t < T:
        seq = seqs[int(torch.randint(0, len(seqs), (1,), generator=g))]
        for tok in seq:
            for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
                if t < T:
                    dom[t] = tok
                    t += 1
            t += int(torch.randint(0, 3, (1,), generator=g))        # blank gap
        t += int(torch.randint(2, 12, (1,), generator=g))
    for t in range(T):
        logits[t, dom[t]] += sharp
        if torch.rand(1, generator=g) < 0.25:                       # competitor above the 0.05 gate
            logits[t, int(torch.randint(0, V, (1,), generator=g))] += sharp - 2.0
    return logits.softmax(1)


def main():
    out = {"V": np.array(V), "tokenset": np.array(sorted(TOKENSET), dtype=np.int32)}
    cases = [(90, 1, 9.0, True), (150, 2, 8.0, True), (60, 3, 10.0, False), (200, 4, 9.0, True), (33, 5, 7.0, True),
             (120, 6, 11.0, True)]
    out["ncases"] = np.array(len(cases))
    for i, (T, seed, sharp, use_set) in enumerate(cases):
        probs = synth_probs(T, seed, sharp)
        hyps = ctc_prefix_beam_search(probs, torch.tensor(T), TOKENSET if use_set else None)
        out[f"probs{i}"] = probs.numpy()
        out[f"use_set{i}"] = np.array(use_set)
        out[f"n{i}"] = np.array(len(hyps))
        out[f"len{i}"] = np.array([len(h[0]) for h in hyps], dtype=np.int32)
        out[f"score{i}"] = np.array([h[1] for h in hyps], dtype=np.float64)
        L = max([len(h[0]) for h in hyps] + [1])
        tok = -np.ones((len(hyps), L), dtype=np.int32)
        frm = -np.ones((len(hyps), L), dtype=np.int32)
        prb = np.zeros((len(hyps), L), dtype=np.float64)
        for k, h in enumerate(hyps):
            assert len(h[0]) == len(h[2])
            for j, (tk, nd) in enumerate(zip(h[0], h[2])):
                assert nd["token"] == tk
                tok[k, j], frm[k, j], prb[k, j] = tk, nd["frame"], nd["prob"]
        out[f"tok{i}"], out[f"frame{i}"], out[f"prob{i}"] = tok, frm, prb
    dst = os.path.join(ROOT, "tests", "golden", "ctc.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, [(int(out[f"n{i}"]), out[f"len{i}"].tolist()[:6]) for i in range(len(cases))])


if __name__ == "__main__":
    main()
