"""Generates tests/golden/*.npz from the REAL reference.  Run in the build container only:

    python oracle/make_golden.py

It imports the reference model code read-only from /root/reference (wekws.model.kws_model,
wekws.utils.cmvn) and torchaudio.compliance.kaldi.fbank (the third-party function the
reference calls, torchaudio 2.11.0 in this image), feeds them deterministic synthetic weights
/ inputs from wekws_b200.synth, and stores inputs + outputs.  The GPU box has no
/root/reference: tests there compare against these files.  TEST INFRASTRUCTURE ONLY.
"""
import io
import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from wekws_b200 import synth                      # noqa: E402
from wekws_b200.configs import model_config      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

from tests.cases import CHUNKS, MODEL_CASES, build_model      # noqa: E402


def build_reference(case):
    from wekws.model.kws_model import init_model
    cfg, model, _ = build_model(case, init_model)
    return cfg, model


def gen_models():
    for case, cfg_name, kw, B in MODEL_CASES:
        cfg, model = build_reference(case)
        idim = cfg["input_dim"]
        has_cmvn = model.global_cmvn is not None
        arrays = {"digest": np.float64(synth.state_digest(model)),
                  "keys": np.array(sorted(model.state_dict().keys()))}
        gru = cfg["backbone"]["type"] == "gru"
        cache = torch.zeros(cfg["backbone"]["num_layers"], B, cfg["hidden_dim"]) if gru else torch.zeros(0, 0, 0)
        xs = []
        with torch.no_grad():
            for i, T in enumerate(CHUNKS):
                x = synth.features(B, T, idim, seed=100 + i, cmvn_like=has_cmvn)
                xs.append(x)
                y, cache = model(x, cache)
                arrays[f"x{i}"] = x.numpy()
                arrays[f"y{i}"] = y.numpy()
                if i in (0, len(CHUNKS) - 1):
                    arrays[f"c{i}"] = cache.numpy()
            # whole utterance in one call == streaming (SURVEY 8a numerical facts)
            full = torch.cat(xs, dim=1)
            c0 = torch.zeros(cfg["backbone"]["num_layers"], B, cfg["hidden_dim"]) if gru else torch.zeros(0, 0, 0)
            yf, cf = model(full, c0)
            arrays["y_full"] = yf.numpy()
            if has_cmvn:
                arrays["cmvn_mean"] = model.global_cmvn.mean.numpy()
                arrays["cmvn_istd"] = model.global_cmvn.istd.numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{case}.npz"), **arrays)
        print(case, "digest", arrays["digest"], "y0 range", float(arrays["y0"].min()), float(arrays["y0"].max()))


def gen_fbank():
    import torchaudio.compliance.kaldi as kaldi
    waves = {}
    pcm = synth.pcm_int16(2, 16000, seed=1234)
    waves["gauss3000_a"] = pcm[0].float()
    waves["gauss3000_b"] = pcm[1].float()
    waves["gauss3_short"] = synth.pcm_int16(1, 4000, seed=5, sigma=3.0)[0].float()
    t = torch.arange(8000, dtype=torch.float64) / 16000.0
    am = 8000.0 * (1.0 + 0.5 * torch.sin(2 * np.pi * 3.0 * t)) * torch.sin(2 * np.pi * 440.0 * t)
    waves["am_tone"] = am.round().clamp(-32768, 32767).float()
    waves["zeros"] = torch.zeros(1200)
    waves["dc"] = torch.full((1000,), 1234.0)
    waves["one_frame"] = synth.pcm_int16(1, 400, seed=9)[0].float()
    waves["too_short"] = synth.pcm_int16(1, 399, seed=10)[0].float()
    waves["chunk_0p3s_carry"] = synth.pcm_int16(1, 4800 + 320, seed=11)[0].float()   # stream_kws_ctc.py:558-563
    arrays = {}
    for name, w in waves.items():
        arrays["wav_" + name] = w.numpy().astype(np.float32)
        for nmel in (80, 40):
            # exactly the reference's call (processor.py:196-202; stream_kws_ctc.py:354-360)
            if w.numel() >= 400:
                f = kaldi.fbank(w.unsqueeze(0), num_mel_bins=nmel, frame_length=25, frame_shift=10, dither=0.0,
                                energy_floor=0.0, sample_frequency=16000)
            else:
                f = torch.empty(0, nmel)
            arrays[f"fbank{nmel}_" + name] = f.numpy()
    # the runtime's window (runtime/core/frontend/fbank.h:90-96 == window_type='hamming')
    arrays["fbank80_hamming_gauss3000_a"] = kaldi.fbank(
        waves["gauss3000_a"].unsqueeze(0), num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
        energy_floor=0.0, sample_frequency=16000, window_type="hamming").numpy()
    # constants the product builds on the host
    arrays["povey_window"] = kaldi._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32).numpy()
    for nmel in (80, 40):
        arrays[f"mel{nmel}"] = kaldi.get_mel_banks(nmel, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "fbank.npz"), **arrays)
    print("fbank cases", len(waves))


def gen_fsmn():
    """FSMN backbone from the live reference: weights, chunked streaming with the 4-D cache, and the
    whole utterance in one call."""
    from wekws.model.kws_model import init_model
    from tests.cases import FSMN_CASES, fsmn_config
    for case in FSMN_CASES:
        cfg = fsmn_config(case)
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(777)
            model = init_model(cfg)
        model.eval()
        arrays = {"sd_" + k: v.numpy() for k, v in model.state_dict().items()}
        cache = torch.zeros(0, 0, 0, 0)
        xs = []
        with torch.no_grad():
            for i, T in enumerate(CHUNKS + (9,)):
                x = synth.features(2, T, cfg["input_dim"], seed=300 + i)
                xs.append(x)
                y, cache = model(x, cache)
                arrays[f"x{i}"], arrays[f"y{i}"], arrays[f"c{i}"] = x.numpy(), y.numpy(), cache.numpy()
            yf, _ = model(torch.cat(xs, dim=1), torch.zeros(0, 0, 0, 0))
            arrays["y_full"] = yf.numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{case}.npz"), **arrays)
        print(case, "cache", tuple(cache.shape))


def gen_mfcc():
    """kaldi.mfcc exactly as wekws/dataset/processor.py:157-166 calls it (the mdtc configs: num_ceps = num_mel_bins = 80)."""
    import torchaudio.compliance.kaldi as kaldi
    fb = np.load(os.path.join(OUT, "fbank.npz"))
    arrays = {}
    for name in ("gauss3000_a", "gauss3_short", "am_tone", "zeros", "one_frame", "chunk_0p3s_carry"):
        w = torch.from_numpy(fb["wav_" + name])
        for nc, nmel in ((80, 80), (40, 40), (13, 23)):
            arrays[f"mfcc{nc}_{nmel}_{name}"] = kaldi.mfcc(
                w.unsqueeze(0), num_ceps=nc, num_mel_bins=nmel, frame_length=25, frame_shift=10, dither=0.0,
                energy_floor=0.0, sample_frequency=16000).numpy()
    arrays["dct80"] = kaldi._get_dct_matrix(80, 80).numpy()
    arrays["dct13_23"] = kaldi._get_dct_matrix(13, 23).numpy()
    arrays["lifter80"] = kaldi._get_lifter_coeffs(80, 22.0).numpy()
    np.savez_compressed(os.path.join(OUT, "mfcc.npz"), **arrays)
    print("mfcc cases", len(arrays))


def gen_cmvn():
    from wekws.utils.cmvn import load_cmvn
    path = synth.write_cmvn_json(80, seed=7)
    ref = load_cmvn(path)
    os.unlink(path)
    np.savez_compressed(os.path.join(OUT, "cmvn.npz"), cmvn=ref)


def gen_init_parity():
    """Same torch seed -> the reference's init_model draws exactly these initial weights."""
    from wekws.model.kws_model import init_model
    arrays = {}
    for name in ("mdtc", "mdtc_small", "ds_tcn", "tcn", "gru"):
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(777)
            m = init_model(model_config(name))
        arrays[name] = np.float64(synth.state_digest(m))
        arrays[name + "_nkeys"] = np.int64(len(m.state_dict()))
    np.savez_compressed(os.path.join(OUT, "init_digest.npz"), **arrays)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_models()
    gen_fbank()
    gen_mfcc()
    gen_fsmn()
    gen_cmvn()
    gen_init_parity()
    print("golden vectors written to", OUT)
