"""Shared description of the golden model cases (used by oracle/make_golden.py and tests)."""
import os

from wekws_b200 import synth
from wekws_b200.configs import model_config

# (case name, config name, kwargs for model_config (+cmvn flag), batch)
MODEL_CASES = [
    ("mdtc", "mdtc", dict(), 2),
    ("mdtc_cmvn_logits", "mdtc", dict(activation="identity", cmvn=True, output_dim=2), 2),
    ("mdtc_small", "mdtc_small", dict(input_dim=40), 2),
    ("ds_tcn", "ds_tcn", dict(), 2),
    ("ds_tcn_ctc", "ds_tcn", dict(activation="identity", output_dim=37, input_dim=40), 1),
    ("tcn", "tcn", dict(), 2),
    ("gru", "gru", dict(), 3),
    ("gru_cmvn", "gru", dict(cmvn=True, input_dim=40, output_dim=2), 2),
]
# FSMN: oracle-only cases -- the product has no FSMN kernel yet.  Small dims (the goldens carry the
# weights), the shipped orders (fsmn_ctc.yaml:42-52) and a strided variant.
FSMN_CASES = {
    "fsmn": dict(left_order=10, right_order=2, left_stride=1, right_stride=1),
    "fsmn_strided": dict(left_order=4, right_order=1, left_stride=2, right_stride=3),   # strides are ignored upstream
}


def fsmn_config(case: str) -> dict:
    bb = dict(type="fsmn", input_affine_dim=24, num_layers=3, linear_dim=36, proj_dim=16, output_affine_dim=20)
    bb.update(FSMN_CASES[case])
    return dict(input_dim=40, output_dim=7, hidden_dim=16, preprocessing=dict(type="none"), backbone=bb,
                classifier=dict(type="identity", dropout=0.1), activation=dict(type="identity"))


CHUNKS = (40, 17, 1)   # streamed back to back, cache carried (17 < pad of the dilation-8 blocks)
CASE_NAMES = [c[0] for c in MODEL_CASES]


def case_config(case: str):
    """-> (cfg dict incl. a temporary cmvn file if needed, batch, cleanup callable)."""
    for name, cfg_name, kw, B in MODEL_CASES:
        if name == case:
            kw = dict(kw)
            cmvn_file = synth.write_cmvn_json(kw.get("input_dim", 80)) if kw.pop("cmvn", False) else None
            cfg = model_config(cfg_name, cmvn_file=cmvn_file, **kw)
            return cfg, B, (lambda: os.unlink(cmvn_file)) if cmvn_file else (lambda: None)
    raise KeyError(case)


def build_model(case: str, factory, seed: int = 777):
    """Instantiates `factory` (reference or wekws_b200 init_model) with the case's synthetic weights."""
    import contextlib
    import io

    import torch
    cfg, B, cleanup = case_config(case)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(seed)
            model = factory(cfg)
    finally:
        cleanup()
    synth.randomize_(model, seed=seed)
    model.eval()
    return cfg, model, B
