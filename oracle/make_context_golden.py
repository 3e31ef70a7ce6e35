#!/usr/bin/env python
"""Golden vectors for the context-expansion / frame-skip transforms (test infrastructure): runs the REFERENCE's own
wekws/dataset/processor.py functions (importable in this container) and stores inputs + outputs in
tests/golden/context.npz.      python oracle/make_context_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import wekws.dataset.processor as P  # noqa: E402

CASES = [(37, 80, 2, 2, 3), (10, 40, 1, 1, 1), (6, 2, 2, 1, 3), (64, 80, 2, 2, 3), (5, 8, 4, 2, 2), (9, 3, 0, 0, 4)]


def main():
    out = {"cases": np.array(CASES, dtype=np.int64)}
    g = torch.Generator().manual_seed(5)
    for i, (T, D, left, right, skip) in enumerate(CASES):
        f = torch.randn(T, D, generator=g)
        y = list(P.frame_skip(P.context_expansion([{"feat": f.clone()}], left, right), skip))[0]["feat"]
        out[f"x{i}"], out[f"y{i}"] = f.numpy(), y.numpy()
    dst = os.path.join(ROOT, "tests", "golden", "context.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, [out[f"y{i}"].shape for i in range(len(CASES))])


if __name__ == "__main__":
    main()
