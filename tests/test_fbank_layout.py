"""CPU check of the index algebra of the Fbank FFT core (wekws_b200/csrc/fbank_core.cuh): the radix 8 / 8 / 4 Stockham
passes with their two exchange layouts and the shuffle-based real-FFT untangle, restated lane by lane in numpy with the
SAME index formulas the kernel uses, must reproduce numpy's rfft -- and the four shared-memory access patterns must be
conflict-free for 8-byte elements (16 distinct bank pairs per half-warp), which is what the header comment claims.
The formulas are read from the source so the test fails if the kernel's layout changes without this restatement."""
import os
import re

import numpy as np

from tests.conftest import ROOT

SRC = open(os.path.join(ROOT, "wekws_b200", "csrc", "fbank_core.cuh")).read()


def _formulas():
    """The four index expressions as they stand in the kernel source."""
    w1 = re.search(r"E\[(8 \* lane \+ \(lane >> 1\) \+ k)\] = make_float2", SRC)
    r1 = re.search(r"const int b = (lane \+ \(lane >> 4\));\s*#pragma unroll\s*for \(int r = 0; r < 8; \+\+r\) \{\s*const float2 v = E\[(b \+ 34 \* r)\]", SRC)
    w2 = re.search(r"const int b = (\(lane & 7\) \+ 72 \* \(lane >> 3\));.*?E\[(b \+ 8 \* k)\] = make_float2", SRC, re.S)
    r2 = re.search(r"v0 = E\[q\], v1 = E\[q \+ 72\], v2 = E\[q \+ 144\], v3 = E\[q \+ 216\]", SRC)
    assert w1 and r1 and w2 and r2, "fbank_core.cuh exchange layout changed: update tests/test_fbank_layout.py"
    assert re.search(r"constexpr int E_SZ = 280;", SRC)
    return (lambda l, k: 8 * l + (l >> 1) + k, lambda l, r: l + (l >> 4) + 34 * r,
            lambda l, k: (l & 7) + 72 * (l >> 3) + 8 * k, lambda q, r: q + 72 * r)


def test_stockham_passes_and_shuffle_untangle_reproduce_rfft():
    w1, r1, w2, r2 = _formulas()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(512)
    x[400:] = 0.0                                           # 400-sample frame zero-padded to 512 (kaldi.py:207-211)
    z = x[0::2] + 1j * x[1::2]                              # even / odd packing: 256-point complex FFT
    W = lambda n, k: np.exp(-2j * np.pi * k / n)
    E = np.full(280, np.nan + 0j)
    for l in range(32):                                     # pass 1: lane owns z[l + 32 i]
        V = np.fft.fft(np.array([z[l + 32 * i] for i in range(8)]))
        for k in range(8):
            assert np.isnan(E[w1(l, k)].real), "two writers for one slot"
            E[w1(l, k)] = V[k] * W(256, l * k)
    regs = [np.array([E[r1(l, r)] for r in range(8)]) for l in range(32)]
    assert not np.isnan(np.array(regs)).any()
    E2 = np.full(280, np.nan + 0j)
    for l in range(32):                                     # pass 2
        V = np.fft.fft(regs[l])
        for k in range(8):
            assert np.isnan(E2[w2(l, k)].real)
            E2[w2(l, k)] = V[k] * W(256, 8 * (l >> 3) * k)
    Z = np.zeros((32, 8), complex)
    for l in range(32):                                     # pass 3: Z[l + 32 i] stays in lane l as element i = hh + 2k
        for hh in range(2):
            v = [E2[r2(l + 32 * hh, j)] for j in range(4)]
            s0, d0, s1, d1 = v[0] + v[2], v[0] - v[2], v[1] + v[3], v[1] - v[3]
            Z[l, hh], Z[l, hh + 2], Z[l, hh + 4], Z[l, hh + 6] = s0 + s1, d0 - 1j * d1, s0 - s1, d0 + 1j * d1
    ref = np.fft.fft(z)
    assert max(abs(Z[l, i] - ref[l + 32 * i]) for l in range(32) for i in range(8)) < 1e-12
    pw = np.zeros(256)
    for l in range(32):                                     # untangle: partner Z[256 - k] comes from lane (32 - l) & 31
        c = Z[(32 - l) & 31]
        for i in range(8):
            k = l + 32 * i
            a, p = Z[l, i], (c[(8 - i) & 7] if l == 0 else c[7 - i])
            sr, si, dr, di = a.real + p.real, a.imag - p.imag, a.real - p.real, a.imag + p.imag
            w = 0.5 * W(512, k)                             # the table holds 0.5 W_512^k
            pp, qq = w.real * dr - w.imag * di, w.real * di + w.imag * dr
            pw[k] = (0.5 * sr + qq) ** 2 + (0.5 * si - pp) ** 2
    assert np.abs(pw - np.abs(np.fft.rfft(x)[:256]) ** 2).max() < 1e-9


def test_exchange_layouts_are_conflict_free_for_8_byte_elements():
    w1, r1, w2, r2 = _formulas()

    def worst(f):
        out = 0
        for half in (range(16), range(16, 32)):
            pairs = [f(l) % 16 for l in half]               # 8-byte bank pair of the element
            out = max(out, max(pairs.count(p) for p in pairs))
        return out
    assert max(worst(lambda l: w1(l, k)) for k in range(8)) == 1
    assert max(worst(lambda l: r1(l, r)) for r in range(8)) == 1
    assert max(worst(lambda l: w2(l, k)) for k in range(8)) == 1
    assert max(worst(lambda l: r2(l + 32 * hh, j)) for hh in range(2) for j in range(4)) == 1
    assert max(max(w1(l, k), w2(l, k)) for l in range(32) for k in range(8)) < 280
