"""Detection statistics on the GPU for max-pooling keyword models.

Replaces the host round trip ``score.py`` (text score file, wekws/bin/score.py:128-137) ->
``compute_det.py`` (threshold sweep with ``window_shift`` skipping, wekws/bin/compute_det.py:76-105)
with one kernel over the posteriors that are already on the device, bit-exact with that pipeline.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native


def det_thresholds(step: float = 0.01) -> torch.Tensor:
    """The thresholds compute_det.py:79-105 visits: ``threshold = 0.0; while threshold <= 1.0: ...; threshold += step``
    accumulated in Python doubles."""
    out, threshold = [], 0.0
    while threshold <= 1.0:
        out.append(threshold)
        threshold += step
    return torch.tensor(out, dtype=torch.float64)


def det_stats(post: torch.Tensor, lengths: Optional[torch.Tensor] = None, step: float = 0.01,
              window_shift: int = 50):
    """post (B, T, K) float32 CUDA posteriors, lengths (B,) valid frames.  Returns (thresholds (n,) float64 on the
    host, max_score (B, K) float64, triggers (B, K, n) int32) -- see include/wekws_b200.h wekws_det_stats."""
    if not post.is_cuda:
        raise RuntimeError("wekws_b200.det_stats runs on CUDA only; got a CPU tensor (no CPU fallback)")
    if post.dim() != 3 or post.dtype != torch.float32:
        raise ValueError("post must be a (B, T, K) float32 tensor")
    post = post.contiguous()
    B, T, K = post.shape
    thr = det_thresholds(step)
    d_thr = thr.to(post.device)
    lens = None if lengths is None else lengths.to(device=post.device, dtype=torch.int32).contiguous()
    max_score = torch.empty(B, K, device=post.device, dtype=torch.float64)
    triggers = torch.empty(B, K, thr.numel(), device=post.device, dtype=torch.int32)
    with torch.cuda.device(post.device):
        rc = _native.lib().wekws_det_stats(
            C.c_void_p(post.data_ptr()), C.c_void_p(lens.data_ptr()) if lens is not None else None, B, T, K,
            C.c_void_p(d_thr.data_ptr()), thr.numel(), int(window_shift), C.c_void_p(max_score.data_ptr()),
            C.c_void_p(triggers.data_ptr()), C.c_void_p(torch.cuda.current_stream(post.device).cuda_stream))
    _native.check(rc, "wekws_det_stats")
    return thr, max_score, triggers


def det_curve(thresholds, max_score, triggers, is_keyword, filler_hours: float, keyword_index: int = 0):
    """Rows (threshold, false_alarm_per_hour, false_reject_rate) exactly as compute_det.py:98-104 writes them, from
    det_stats outputs: is_keyword (B,) bool marks the utterances whose transcript is the keyword."""
    is_keyword = torch.as_tensor(is_keyword, dtype=torch.bool).cpu()
    ms = max_score[:, keyword_index].double().cpu()
    tr = triggers[:, keyword_index].cpu()
    nk = int(is_keyword.sum())
    rows = []
    for i, th in enumerate(thresholds.tolist()):
        num_false_reject = int((ms[is_keyword] < th).sum())
        num_false_alarm = max(int(tr[~is_keyword, i].sum()), 1e-6)
        frr = num_false_reject / nk if nk else 0.0
        fa = num_false_alarm / filler_hours if filler_hours else 0.0
        rows.append((th, fa, frr))
    return rows


def context_expansion(feats: torch.Tensor, left: int = 1, right: int = 1, skip_rate: int = 1,
                      lengths: Optional[torch.Tensor] = None):
    """wekws/dataset/processor.py:267-312 (context_expansion then frame_skip) on the device, batched:
    feats (B, T, D) float32 CUDA [, lengths (B,)] -> (expanded (B, m, D*(left+right+1)), new_lengths (B,) int32) with
    m = ceil((T - right) / skip_rate); stream b keeps ceil((lengths[b] - right) / skip_rate) rows, the rest are zero
    (the zero padding pad_sequence would add)."""
    if not feats.is_cuda:
        raise RuntimeError("wekws_b200.context_expansion runs on CUDA only; got a CPU tensor (no CPU fallback)")
    if feats.dim() != 3 or feats.dtype != torch.float32:
        raise ValueError("feats must be a (B, T, D) float32 tensor")
    feats = feats.contiguous()
    B, T, D = feats.shape
    lib = _native.lib()
    m = int(lib.wekws_context_expand_frames(T, int(right), int(skip_rate)))
    out = torch.empty(B, m, D * (left + right + 1), device=feats.device, dtype=torch.float32)
    lens = None if lengths is None else lengths.to(device=feats.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(feats.device):
        rc = lib.wekws_context_expand(
            C.c_void_p(feats.data_ptr()), C.c_void_p(lens.data_ptr()) if lens is not None else None, B, T, D, int(left),
            int(right), int(skip_rate), C.c_void_p(out.data_ptr()), m,
            C.c_void_p(torch.cuda.current_stream(feats.device).cuda_stream))
    _native.check(rc, "wekws_context_expand")
    n = torch.full((B,), T, dtype=torch.int64) if lengths is None else lengths.detach().cpu().to(torch.int64)
    new_len = torch.where(n > right, (n - right + skip_rate - 1) // skip_rate, torch.zeros_like(n)).to(torch.int32)
    return out, new_len
