"""Parity of the CUDA path (through the C-ABI) against the golden vectors of the real reference
and against the CPU oracle.  Tolerances: posteriors <= 1e-4 max-abs (north_star); caches <= 1e-4
scaled by magnitude; log-mel features <= 1e-3 max-abs / 1e-5 mean-abs (the oracle's own
fp32-vs-fp64 noise floor is 9e-5..4.3e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import kws_oracle as O
from tests.cases import CASE_NAMES, CHUNKS, build_model
from tests.conftest import ROOT, golden
from wekws_b200 import Fbank, init_model, model_config, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_POST = 1e-4
TOL_FEAT_MAX, TOL_FEAT_MEAN = 1e-3, 1e-5


def _tol(ref):
    return TOL_POST * max(1.0, float(np.abs(ref).max()))


@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(case):
        if case not in cache:
            cfg, m, B = build_model(case, init_model)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            cache[case] = (cfg, m.to(DEV), sd, B)
        return cache[case]
    return get


def _zero_cache(cfg, m, B, dev="cpu"):
    if cfg["backbone"]["type"] == "gru":
        return torch.zeros(cfg["backbone"]["num_layers"], B, cfg["hidden_dim"], device=dev)
    return torch.zeros(0, 0, 0)


@pytest.mark.parametrize("case", CASE_NAMES)
def test_streaming_matches_reference_golden(case, models):
    g = golden("model_" + case)
    cfg, m, sd, B = models(case)
    assert abs(synth.state_digest(m) - float(g["digest"])) < 1e-6 * float(g["digest"])
    cache = _zero_cache(cfg, m, B, DEV)
    for i, T in enumerate(CHUNKS):
        y, cache = m(torch.from_numpy(g[f"x{i}"]).to(DEV), cache)
        assert y.shape == g[f"y{i}"].shape
        err = np.abs(y.cpu().numpy() - g[f"y{i}"]).max()
        assert err <= _tol(g[f"y{i}"]), (case, i, err)
        if f"c{i}" in g:
            cerr = np.abs(cache.cpu().numpy() - g[f"c{i}"]).max()
            assert cerr <= _tol(g[f"c{i}"]), (case, i, cerr)
    full = torch.cat([torch.from_numpy(g[f"x{i}"]) for i in range(len(CHUNKS))], dim=1).to(DEV)
    yf, _ = m(full, _zero_cache(cfg, m, B, DEV))
    assert np.abs(yf.cpu().numpy() - g["y_full"]).max() <= _tol(g["y_full"])


@pytest.mark.parametrize("case", CASE_NAMES)
@pytest.mark.parametrize("B,T", [(1, 1), (5, 7), (3, 33), (7, 40), (2, 131)])
def test_matches_oracle_with_random_cache(case, B, T, models):
    cfg, m, sd, _ = models(case)
    gen = torch.Generator().manual_seed(B * 1000 + T)
    x = synth.features(B, T, cfg["input_dim"], seed=B * 31 + T, cmvn_like=m.global_cmvn is not None)
    if cfg["backbone"]["type"] == "gru":
        cache = torch.randn(cfg["backbone"]["num_layers"], B, cfg["hidden_dim"], generator=gen)
    else:
        cache = torch.randn(B, m.hdim, m.backbone.padding, generator=gen)
    y_ref, c_ref = O.kws_forward(sd, cfg, x, cache)
    y, c = m(x.to(DEV), cache.to(DEV))
    assert (y.cpu() - y_ref).abs().max() <= _tol(y_ref.numpy())
    assert (c.cpu() - c_ref).abs().max() <= _tol(c_ref.numpy())


def test_empty_cache_equals_zero_cache_and_forward_softmax(models):
    cfg, m, sd, _ = models("mdtc_cmvn_logits")
    x = synth.features(4, 25, 80, seed=8, cmvn_like=True).to(DEV)
    y0, c0 = m(x)                                            # default arg: CPU zeros(0,0,0) (kws_model.py:68)
    y1, c1 = m(x, torch.zeros(4, 64, 244, device=DEV))
    assert torch.equal(y0, y1) and torch.equal(c0, c1)
    ys, _ = m.forward_softmax(x)
    assert (ys - y0.softmax(2)).abs().max() <= 1e-6
    y_ref, _ = O.kws_forward(sd, cfg, x.cpu(), None, softmax=True)
    assert (ys.cpu() - y_ref).abs().max() <= TOL_POST


def test_edge_shapes_and_interface(models):
    cfg, m, sd, _ = models("mdtc")
    y, c = m(torch.zeros(0, 5, 80, device=DEV))
    assert y.shape == (0, 5, 1) and c.shape == (0, 64, 244)
    y, c = m(torch.zeros(2, 0, 80, device=DEV))
    assert y.shape == (2, 0, 1) and c.shape == (2, 64, 244) and float(c.abs().max()) == 0.0
    x = synth.features(2, 9, 80, seed=1).to(DEV)
    xt = x.transpose(0, 1).contiguous().transpose(0, 1)      # non-contiguous view, same values
    assert not xt.is_contiguous()
    assert torch.equal(m(xt)[0], m(x)[0])
    with pytest.raises(ValueError):
        m(x, torch.zeros(2, 64, 100, device=DEV))
    with pytest.raises(TypeError):
        m(x.double())
    # returned cache is a fresh tensor, input cache untouched
    cin = torch.randn(2, 64, 244, device=DEV)
    keep = cin.clone()
    _, cout = m(x, cin)
    assert torch.equal(cin, keep) and cout.data_ptr() != cin.data_ptr()
    # weights reloaded -> native pack refreshed
    y_before = m(x)[0]
    sd2 = {k: (v * 1.01 if v.dtype.is_floating_point and "running_var" not in k else v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd2)
    assert not torch.equal(m(x)[0], y_before)
    m.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    assert torch.equal(m(x)[0], y_before)


@pytest.mark.parametrize("name", ["mdtc", "tcn", "ds_tcn", "gru"])
def test_full_size_batch_properties(name, models):
    """BASELINE sizes (B=1024 streams x 40 frames): streaming == full utterance, rows independent."""
    cfg, m, sd, _ = models({"mdtc": "mdtc", "tcn": "tcn", "ds_tcn": "ds_tcn", "gru": "gru"}[name])
    B = 1024 if name != "gru" else 512
    x = synth.features(B, 80, 80, seed=99).to(DEV)
    c0 = _zero_cache(cfg, m, B, DEV)
    y_full, c_full = m(x, c0)
    ya, ca = m(x[:, :40].contiguous(), c0)
    yb, cb = m(x[:, 40:].contiguous(), ca)
    assert (torch.cat((ya, yb), 1) - y_full).abs().max() <= 2e-5
    assert (cb - c_full).abs().max() <= 2e-4
    # a permutation of the streams permutes the outputs (no cross-stream leakage, any tile shape)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(DEV)
    yp, _ = m(x[perm].contiguous(), c0)
    assert torch.equal(yp, y_full[perm])
    # spot-check rows against the oracle
    rows = [0, 1, 511, B - 1]
    y_ref, _ = O.kws_forward(sd, cfg, x[rows].cpu(), _zero_cache(cfg, m, len(rows)) if name == "gru" else None)
    assert (y_full[rows].cpu() - y_ref).abs().max() <= TOL_POST


def _check_feats(out, ref, what, wav=None, **kw):
    """Direct agreement with the reference's fp32 output within the SURVEY 8c gate, OR -- for
    signals where the reference's own fp32 rounding exceeds that gate (pure tones: energy in a
    few bins, the rest is rounding noise) -- at least as close to the float64 evaluation of the
    same formulas as the reference itself is (x1.5 slack)."""
    assert out.shape == ref.shape, what
    if not ref.size:
        return
    d = np.abs(out - ref)
    if d.max() <= TOL_FEAT_MAX and d.mean() <= TOL_FEAT_MEAN:
        return
    assert wav is not None, (what, d.max(), d.mean())
    truth = O.fbank(wav, dtype=torch.float64, **kw).numpy()
    e_ref, e_out = np.abs(ref - truth), np.abs(out - truth)
    assert e_out.max() <= max(TOL_FEAT_MAX, 1.5 * e_ref.max()), (what, e_out.max(), e_ref.max())
    assert e_out.mean() <= max(TOL_FEAT_MEAN, 1.5 * e_ref.mean()), (what, e_out.mean(), e_ref.mean())


def test_fbank_matches_reference_golden():
    g = golden("fbank")
    names = [k[4:] for k in g.files if k.startswith("wav_")]
    for nmel in (80, 40):
        fb = Fbank(nmel)
        for name in names:
            wav = torch.from_numpy(g["wav_" + name])
            ref = g[f"fbank{nmel}_{name}"]
            _check_feats(fb(wav.to(DEV)).cpu().numpy(), ref, (name, nmel, "f32"), wav, num_mel_bins=nmel)
            _check_feats(fb(wav.to(torch.int16).to(DEV)).cpu().numpy(), ref, (name, nmel, "s16"), wav,
                         num_mel_bins=nmel)
    wav = torch.from_numpy(g["wav_gauss3000_a"])
    ham = Fbank(80, window_type="hamming")(wav.to(DEV)).cpu().numpy()
    _check_feats(ham, g["fbank80_hamming_gauss3000_a"], "hamming", wav, window_type="hamming")
    z = Fbank(80)(torch.zeros(2, 1200, device=DEV))
    assert torch.all(z == float(np.log(np.float32(O.EPS))))


def _check_mfcc(out, ref, what, wav, nc, nmel):
    """The reference's own fp32 MFCC sits 3.5e-3 max / 3.5e-4 mean from the float64 evaluation (sgemm over 80
    log-mels of magnitude ~16); the kernel must agree with it to that order, or be at least as close to float64."""
    assert out.shape == ref.shape, what
    if not ref.size:
        return
    d = np.abs(out - ref)
    if d.max() <= 6e-3 and d.mean() <= 6e-4:
        return
    truth = O.mfcc(wav, nc, nmel, dtype=torch.float64).numpy()
    e_ref, e_out = np.abs(ref - truth), np.abs(out - truth)
    assert e_out.max() <= max(6e-3, 1.5 * e_ref.max()), (what, e_out.max(), e_ref.max())
    assert e_out.mean() <= max(6e-4, 1.5 * e_ref.mean()), (what, e_out.mean(), e_ref.mean())


def test_mfcc_matches_reference_golden():
    """kaldi.mfcc as processor.py:157-166 calls it; fused DCT + lifter epilogue of the Fbank kernel."""
    from wekws_b200 import Mfcc, mfcc
    g, fbg = golden("mfcc"), golden("fbank")
    fes = {}
    for k in [k for k in g.files if k.startswith("mfcc")]:
        head, nmel, name = k.split("_", 2)
        nc, nmel = int(head[4:]), int(nmel)
        fe = fes.setdefault((nc, nmel), Mfcc(nc, nmel))
        wav = torch.from_numpy(fbg["wav_" + name])
        _check_mfcc(fe(wav.to(DEV)).cpu().numpy(), g[k], (k, "f32"), wav, nc, nmel)
        _check_mfcc(fe(wav.to(torch.int16).to(DEV)).cpu().numpy(), g[k], (k, "s16"), wav, nc, nmel)
    # the functional form with the reference's call signature, (1, N) -> (m, num_ceps)
    wav = torch.from_numpy(fbg["wav_gauss3000_a"])
    f = mfcc(wav.unsqueeze(0).to(DEV), num_ceps=80, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
             energy_floor=0.0, sample_frequency=16000)
    _check_mfcc(f.cpu().numpy(), g["mfcc80_80_gauss3000_a"], "functional", wav, 80, 80)
    assert Mfcc(80, 80)(torch.zeros(399, device=DEV)).shape == (0, 80)
    with pytest.raises(AssertionError):
        Mfcc(81, 80)


def test_mfcc_batched_ragged_cmvn_and_model_pipeline(models):
    from wekws_b200 import Mfcc
    fe = Mfcc(80, 80)
    pcm = synth.pcm_int16(5, 16000, seed=22)
    lens = torch.tensor([16000, 15999, 8000, 400, 399], dtype=torch.int32)
    mean, istd = torch.randn(80) * 3, torch.rand(80) * 0.2 + 0.05
    out = fe(pcm.to(DEV), lengths=lens.to(DEV), mean=mean.to(DEV), istd=istd.to(DEV)).cpu()
    assert out.shape == (5, 98, 80)
    for b in range(5):
        ref = O.mfcc(pcm[b, :lens[b]].float())
        n = ref.shape[0]
        if n:
            d = (out[b, :n] - O.global_cmvn(ref, mean, istd)).abs()
            assert d.max() <= 6e-3 and d.mean() <= 6e-4
        assert float(out[b, n:].abs().max() if n < 98 else 0.0) == 0.0
    # MFCC features into the model (the shipped mdtc recipe): posteriors against the oracle chain
    cfg, m, sd, _ = models("mdtc_cmvn_logits")
    y, _ = m(fe(pcm.to(DEV)))
    ref_f = torch.stack([O.mfcc(pcm[b].float()) for b in range(5)])
    y_ref, _ = O.kws_forward(sd, cfg, ref_f, None)
    # the reference's own fp32-vs-float64 MFCC noise (3.5e-3) moves these logits (|y| ~ 13) by 4.5e-4 [measured with
    # the oracle]; two fp32 implementations can differ by twice that, hence 3x the posterior gate for this chain
    assert (y.cpu() - y_ref).abs().max() <= 3 * _tol(y_ref.numpy())


def test_fbank_batched_ragged_cmvn_and_chunked_streaming():
    fb = Fbank(80)
    pcm = synth.pcm_int16(6, 16000, seed=21)
    lens = torch.tensor([16000, 15999, 8000, 400, 399, 12345], dtype=torch.int32)
    mean = torch.randn(80) + 15
    istd = torch.rand(80) + 0.2
    out = fb(pcm.to(DEV), lengths=lens.to(DEV), mean=mean.to(DEV), istd=istd.to(DEV)).cpu()
    assert out.shape == (6, 98, 80)
    for b in range(6):
        ref = O.fbank(pcm[b, :lens[b]].float())
        n = ref.shape[0]
        if n:
            refn = O.global_cmvn(ref, mean, istd)
            d = (out[b, :n] - refn).abs()
            assert d.max() <= TOL_FEAT_MAX and d.mean() <= TOL_FEAT_MEAN
        assert float(out[b, n:].abs().max() if n < 98 else 0.0) == 0.0     # zero padded like pad_sequence
    # chunked Fbank with the 320-sample carry == whole-utterance Fbank (stream_kws_ctc.py:347-364)
    wav = synth.pcm_int16(1, 16000 * 2, seed=3)[0].to(DEV)
    whole = fb(wav)
    parts, rem = [], wav[:0]
    for s in range(0, wav.numel(), 4800):
        buf = torch.cat((rem, wav[s:s + 4800]))
        f = fb(buf)
        parts.append(f)
        rem = buf[f.shape[0] * 160:]
    assert torch.equal(torch.cat(parts), whole)


def test_pcm_to_posterior_pipeline_matches_oracle(models):
    cfg, m, sd, _ = models("mdtc_cmvn_logits")
    fb = Fbank(80)
    pcm = synth.pcm_int16(5, 16000, seed=2)
    feats = fb(pcm.to(DEV))
    y, c = m(feats)
    ref_f = torch.stack([O.fbank(pcm[b].float()) for b in range(5)])
    y_ref, c_ref = O.kws_forward(sd, cfg, ref_f, None)
    assert (y.cpu() - y_ref).abs().max() <= _tol(y_ref.numpy())
    assert (c.cpu() - c_ref).abs().max() <= 2e-3 * max(1.0, float(c_ref.abs().max()))


@pytest.mark.parametrize("case,batch", [("mdtc", 40), ("ds_tcn", 17), ("gru", 1)])
def test_native_runtime_shim_streams_like_the_python_model(case, batch, models, tmp_path):
    """wekws::KeywordSpotting (C++ over the C ABI, no Python/ONNX) fed batch by batch as kws_main.cc:43-61 does ==
    KWSModel streamed with the same chunking; Reset() starts a new stream."""
    import subprocess
    from wekws_b200 import export_native
    cfg, m, sd, _ = models(case)
    export_native(m, str(tmp_path / "m.wkb"))
    T = 100
    x = synth.features(1, T, 80, seed=9, cmvn_like=m.global_cmvn is not None)
    x[0].numpy().tofile(str(tmp_path / "feats.f32"))
    exe = os.path.join(ROOT, "wekws_b200", "runtime", "kws_main_b200")
    reset_every = 3
    r = subprocess.run([exe, str(tmp_path / "m.wkb"), str(tmp_path / "feats.f32"), "80", str(batch), str(reset_every)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert f"cache_dim: {m.hdim}" in r.stdout
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("frame ")]
    assert [int(p[1]) for p in rows] == list(range(T))
    got = torch.tensor([[float(v) for v in p[3:]] for p in rows])
    gru = case == "gru"
    ref, cache, nb = [], None, 0
    for s0 in range(0, T, batch):
        if cache is None:
            cache = torch.zeros(cfg["backbone"]["num_layers"], 1, cfg["hidden_dim"], device=DEV) if gru else torch.zeros(0, 0, 0)
        y, cache = m(x[:, s0:s0 + batch].to(DEV), cache)
        ref.append(y[0].cpu())
        nb += 1
        if nb % reset_every == 0:
            cache = None
    ref = torch.cat(ref)
    assert got.shape == ref.shape
    assert (got - ref).abs().max() <= 2e-6          # same kernels, same chunking: only the 9-digit text round trip
    # ... and against the CPU oracle streamed with the same chunking and resets (the checker proper)
    oref, ocache, nb = [], None, 0
    for s0 in range(0, T, batch):
        if ocache is None and gru:
            ocache = torch.zeros(cfg["backbone"]["num_layers"], 1, cfg["hidden_dim"])
        y, ocache = O.kws_forward(sd, cfg, x[:, s0:s0 + batch], ocache)
        oref.append(y[0])
        nb += 1
        if nb % reset_every == 0:
            ocache = None
    oref = torch.cat(oref)
    assert (got - oref).abs().max() <= _tol(oref.numpy())


def test_native_runtime_pcm_to_posterior_like_kws_main(models, tmp_path):
    """The whole kws_main.cc flow in C++: int16 PCM -> wenet::FeaturePipeline (GPU Fbank, Hamming window as the
    runtime uses, fed in 0.3 s chunks from a producer thread) -> wekws::KeywordSpotting, against the Python chain."""
    import subprocess
    from wekws_b200 import export_native
    cfg, m, sd, _ = models("mdtc")
    export_native(m, str(tmp_path / "m.wkb"))
    pcm = synth.pcm_int16(1, 16000 * 2 + 123, seed=13)[0]
    pcm.numpy().astype("<i2").tofile(str(tmp_path / "a.s16"))
    exe = os.path.join(ROOT, "wekws_b200", "runtime", "kws_main_b200")
    batch = 32
    r = subprocess.run([exe, "--pcm", str(tmp_path / "m.wkb"), str(tmp_path / "a.s16"), "80", str(batch), "4800"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("frame ")]
    got = torch.tensor([[float(v) for v in p[3:]] for p in rows])
    feats = Fbank(80, window_type="hamming")(pcm.to(DEV)).unsqueeze(0)
    T = feats.shape[1]
    assert T == 1 + (pcm.numel() - 400) // 160 and [int(p[1]) for p in rows] == list(range(T))
    ref, cache = [], torch.zeros(0, 0, 0)
    for s0 in range(0, T, batch):
        y, cache = m(feats[:, s0:s0 + batch], cache)
        ref.append(y[0].cpu())
    ref = torch.cat(ref)
    # the C++ window / mel tables are built in double, the Python ones with torch fp32 ops: the features differ in
    # the last bits and the posteriors by 2.4e-5 (measured) -- the posterior gate applies
    assert (got - ref).abs().max() <= TOL_POST
    # ... and against the CPU oracle: Hamming-window Fbank (pinned to the compiled reference front-end by
    # tests/test_compiled_reference.py) -> reference forward with the same 32-frame batches
    ofeats = O.fbank(pcm.float(), window_type="hamming").unsqueeze(0)
    oref, ocache = [], None
    for s0 in range(0, T, batch):
        y, ocache = O.kws_forward(sd, cfg, ofeats[:, s0:s0 + batch], ocache)
        oref.append(y[0])
    assert (got - torch.cat(oref)).abs().max() <= TOL_POST


def test_det_stats_bit_exact_with_score_file_pipeline():
    """the compute_det.py threshold sweep on the device == the host pipeline through the '{:.6f}' score
    file, integer counts and rounded maxima bit for bit (oracle: oracle/kws_oracle.py det_stats)."""
    from wekws_b200 import det_curve, det_stats
    g = torch.Generator().manual_seed(11)
    B, T, K = 7, 230, 2
    post = torch.sigmoid(torch.randn(B, T, K, generator=g) * 3)
    post[0, :5, 0] = torch.tensor([0.5, 0.4999995, 0.5000005, 0.01, 1.0])     # text-rounding ties and exact thresholds
    post[1] = 0.0
    lens = torch.tensor([230, 229, 100, 51, 1, 0, 230], dtype=torch.int32)
    for ws in (50, 1, 7):
        thr, ms, tr = det_stats(post.to(DEV), lens.to(DEV), step=0.01, window_shift=ws)
        o_thr, o_ms, o_tr = O.det_stats(post, lens, 0.01, ws)
        assert thr.tolist() == o_thr and len(o_thr) in (100, 101)
        assert tr.cpu().tolist() == o_tr, ws
        got = ms.cpu()
        assert got.dtype == torch.float64          # compared in double like compute_det.py:84
        for b in range(B):
            for k in range(K):
                assert got[b, k].item() == o_ms[b][k], (b, k)
    thr, ms, tr = det_stats(post.to(DEV), None, window_shift=50)
    assert tr.cpu().tolist() == O.det_stats(post, None, 0.01, 50)[2]
    rows = det_curve(thr, ms, tr, [True, False, True, False, False, False, True], filler_hours=0.5)
    assert len(rows) == thr.numel() and rows[0][2] == 0.0 and rows[-1][1] >= 0.0
    with pytest.raises(RuntimeError):
        det_stats(post)


def test_det_stats_reproduces_reference_compute_det_files():
    """Device sweep + det_curve == the stats files written by the reference's own compute_det.py (golden made by
    oracle/make_det_golden.py): thresholds, false alarms per hour and false-reject rates, text for text."""
    from wekws_b200 import det_curve, det_stats
    g = golden("det_stats")
    post, lens = torch.from_numpy(g["post"]).to(DEV), torch.from_numpy(g["lens"]).to(DEV)
    kinds, durs = g["kinds"].tolist(), g["durations"].tolist()
    for si, (ws, step) in enumerate(g["settings"].tolist()):
        thr, ms, tr = det_stats(post, lens, step=step, window_shift=int(ws))
        for k in range(post.shape[2]):
            filler = 0.0
            for b, kd in enumerate(kinds):
                if kd != k:
                    filler += durs[b]
            rows = det_curve(thr, ms, tr, [kd == k for kd in kinds], filler_hours=filler / 3600.0, keyword_index=k)
            text = "".join("{:.6f} {:.6f} {:.6f}\n".format(*r) for r in rows)
            assert text == str(g[f"stats_{si}_{k}"]), (si, k)


def test_launch_counter_counts_our_kernels(native, models):
    cfg, m, sd, _ = models("mdtc")
    x = synth.features(2, 40, 80, seed=1).to(DEV)
    m(x)
    n0 = native.launch_count()
    m(x)
    Fbank(80)(torch.zeros(1, 16000, device=DEV))
    assert native.launch_count() == n0 + 2


@pytest.mark.parametrize("case", ["mdtc", "mdtc_cmvn_logits", "tcn", "ds_tcn"])
def test_tensor_core_and_fp32_paths_agree(case, models):
    """mdtc, dense tcn (hidden 64) and ds_tcn (hidden 256) run on tcgen05 (bf16x3) by default; the FP32-FMA kernel
    is the exact path."""
    cfg, m, sd, _ = models(case)
    B, T = 37, 40
    x = synth.features(B, T, 80, seed=5, cmvn_like=m.global_cmvn is not None).to(DEV)
    cache = torch.randn(B, m.hdim, m.backbone.padding, generator=torch.Generator().manual_seed(3)).to(DEV)
    try:
        m.precision = "fp32"
        y32, c32 = m(x, cache)
        assert not m.uses_tensor_cores(T)
        m.precision = "auto"
        ytc, ctc = m(x, cache)
        assert m.uses_tensor_cores(T) and not m.uses_tensor_cores(4)
    finally:
        m.precision = "auto"
    assert (ytc - y32).abs().max() <= 0.5 * _tol(y32.cpu().numpy())
    assert (ctc - c32).abs().max() <= 0.5 * _tol(c32.cpu().numpy())
    y_ref, c_ref = O.kws_forward(sd, cfg, x.cpu(), cache.cpu())
    assert (ytc.cpu() - y_ref).abs().max() <= _tol(y_ref.numpy())
    assert (y32.cpu() - y_ref).abs().max() <= 0.2 * _tol(y_ref.numpy())
