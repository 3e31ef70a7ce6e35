"""Deterministic synthetic weights and inputs: there is no network
for datasets or checkpoints, so tests, golden vectors and bench.py all use these."""
from __future__ import annotations

import json
import os
import tempfile

import torch


def randomize_(model: torch.nn.Module, seed: int = 777) -> torch.nn.Module:
    """Fills every state_dict entry (sorted by key, own Generator) so that eval-mode
    BatchNorm and CMVN are non-trivial and activations stay O(1) through all blocks.
    Works identically on the reference model and on wekws_b200.KWSModel (same keys)."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name in sorted(sd):
            t = sd[name]
            if name.endswith("num_batches_tracked"):
                t.fill_(100)
                continue
            cpu = torch.empty(t.shape, dtype=torch.float32)
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_mean":
                cpu.normal_(0.0, 0.5, generator=g)
            elif leaf == "running_var":
                cpu.uniform_(0.5, 2.0, generator=g)
            elif name == "global_cmvn.mean":
                cpu.normal_(15.0, 2.0, generator=g)
            elif name == "global_cmvn.istd":
                cpu.uniform_(0.2, 0.5, generator=g)
            elif leaf == "weight" and t.dim() == 1:          # BatchNorm gamma
                cpu.uniform_(0.5, 1.5, generator=g)
            elif leaf.startswith("bias") and t.dim() == 1:
                cpu.normal_(0.0, 0.1, generator=g)
            else:                                            # conv / linear / GRU matrices
                fan_in = t[0].numel() if t.dim() > 1 else t.numel()
                bound = 1.0 / max(fan_in, 1) ** 0.5
                cpu.uniform_(-bound, bound, generator=g)
            t.copy_(cpu)
    return model


def state_digest(model: torch.nn.Module) -> float:
    """Cheap fingerprint used to make sure two processes generated the same weights."""
    s = 0.0
    for name, t in sorted(model.state_dict().items()):
        if t.dtype.is_floating_point:
            s += float(t.double().abs().sum())
    return s


def write_cmvn_json(dim: int, seed: int = 7, path: str | None = None) -> str:
    """A cmvn stats file in the schema tools/compute_cmvn_stats.py:144-148 writes."""
    g = torch.Generator().manual_seed(seed)
    n = 1000
    mean = torch.empty(dim).normal_(15.0, 2.0, generator=g)
    std = torch.empty(dim).uniform_(2.0, 5.0, generator=g)
    stats = {"mean_stat": (mean * n).tolist(), "var_stat": ((std * std + mean * mean) * n).tolist(), "frame_num": n}
    if path is None:
        fd, path = tempfile.mkstemp(suffix="_cmvn.json")
        os.close(fd)
    with open(path, "w") as f:
        json.dump(stats, f)
    return path


def pcm_int16(batch: int, num_samples: int, seed: int = 1234, sigma: float = 3000.0) -> torch.Tensor:
    """int16(clip(N(0, sigma^2))) synthetic 16 kHz audio."""
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(batch, num_samples).normal_(0.0, sigma, generator=g)
    return x.clamp_(-32768, 32767).round_().to(torch.int16)


def features(batch: int, frames: int, dim: int, seed: int = 4321, cmvn_like: bool = False) -> torch.Tensor:
    """N(0,1) features, or log-mel-like N(15, 3^2) when the model carries a CMVN."""
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(batch, frames, dim).normal_(0.0, 1.0, generator=g)
    return x * 3.0 + 15.0 if cmvn_like else x
