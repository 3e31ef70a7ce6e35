"""Round-2 parity cases (VERDICT r01 "What's weak" 1-4): BASELINE.json shapes that had no test, absolute
1e-4 gates on SIGMOID posteriors over >= 1000 clips, the hidden-32 tile bug, stale weight packs, patch_reference().
Everything goes through the C ABI; the checker is oracle/kws_oracle.py (pinned to the reference's goldens)."""
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import kws_oracle as O
from tests.cases import build_model
from wekws_b200 import Fbank, Mfcc, init_model, model_config, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_POST = 1e-4          # north_star: <= 1e-4 max-abs on the posterior scores, ABSOLUTE (no magnitude scaling here)


def _model(name, seed=777, **kw):
    cmvn = kw.pop("cmvn", False)
    cmvn_file = synth.write_cmvn_json(kw.get("input_dim", 80)) if cmvn else None
    try:
        cfg = model_config(name, cmvn_file=cmvn_file, **kw)
        torch.manual_seed(seed)
        m = synth.randomize_(init_model(cfg), seed=seed).eval()
    finally:
        if cmvn_file:
            os.unlink(cmvn_file)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return cfg, m.to(DEV), sd


def test_pcm_to_sigmoid_posterior_1250_clips():
    """BASELINE configs[4] per-GPU shape: 1250 one-second clips, raw int16 PCM -> Fbank -> CMVN -> mdtc -> sigmoid.
    Max-abs error over all 1250 x 98 posteriors against the CPU oracle (torchaudio-equivalent Fbank + reference
    forward) <= 1e-4 absolute."""
    cfg, m, sd = _model("mdtc", cmvn=True)
    pcm = synth.pcm_int16(1250, 16000, seed=1234)
    y, c = m(Fbank(80)(pcm.to(DEV)))
    ref_f = torch.stack([O.fbank(pcm[b].float()) for b in range(pcm.shape[0])])
    y_ref, c_ref = O.kws_forward(sd, cfg, ref_f, None)
    err = float((y.cpu() - y_ref).abs().max())
    cerr = float(((c.cpu() - c_ref).abs() / c_ref.abs().clamp_min(1.0)).max())
    print(f"pcm->posterior 1250 clips (cmvn): posterior max-abs {err:.3e}, cache rel {cerr:.3e}, "
          f"posterior range [{float(y_ref.min()):.3f}, {float(y_ref.max()):.3f}]")
    assert y.shape == (1250, 98, 1) and err <= TOL_POST
    assert cerr <= 2e-3


def test_pcm_to_sigmoid_posterior_1250_clips_without_cmvn():
    """Same without CMVN: raw log-mel values (10..20) enter the first Linear unscaled, posteriors span [0.001, 1].  Here
    the fp32 reference's OWN distance to the exact (float64) evaluation of the same formulas reaches the 1e-4 bar
    (its fp32 FFT noise in low-energy bins), so the gate is: within 1e-4 of the fp32 reference, or at least
    as close to the float64 evaluation as the fp32 reference itself is (x1.25).  All three distances are printed."""
    cfg, m, sd = _model("mdtc", cmvn=False)
    pcm = synth.pcm_int16(1250, 16000, seed=1234)
    y, _ = m(Fbank(80)(pcm.to(DEV)))
    ref_f = torch.stack([O.fbank(pcm[b].float()) for b in range(pcm.shape[0])])
    y_ref, _ = O.kws_forward(sd, cfg, ref_f, None)
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    f64 = torch.stack([O.fbank(pcm[b].double(), dtype=torch.float64) for b in range(pcm.shape[0])])
    y64, _ = O.kws_forward(sd64, cfg, f64, None)
    e_ref = float((y.cpu() - y_ref).abs().max())
    e_ours64 = float((y.cpu().double() - y64).abs().max())
    e_ref64 = float((y_ref.double() - y64).abs().max())
    print(f"pcm->posterior 1250 clips (no cmvn): ours vs fp32 reference {e_ref:.3e}; vs float64: ours {e_ours64:.3e}, "
          f"fp32 reference {e_ref64:.3e}")
    assert e_ref <= TOL_POST or e_ours64 <= max(TOL_POST, 1.25 * e_ref64)
    assert e_ref <= 2.0 * TOL_POST


def test_mfcc_to_sigmoid_posterior_1000_clips():
    """The shipped mdtc front-end (mdtc.yaml:8-14: mfcc, 80 ceps) -> mdtc -> sigmoid over 1000 clips, absolute gate."""
    cfg, m, sd = _model("mdtc", cmvn=True)
    pcm = synth.pcm_int16(1000, 16000, seed=99)
    y, _ = m(Mfcc(80, 80)(pcm.to(DEV)))
    ref_f = torch.stack([O.mfcc(pcm[b].float(), 80, 80) for b in range(pcm.shape[0])])
    y_ref, _ = O.kws_forward(sd, cfg, ref_f, None)
    err = float((y.cpu() - y_ref).abs().max())
    print(f"mfcc->posterior 1000 clips: posterior max-abs {err:.3e}")
    assert err <= TOL_POST


def test_gru_b512_t1_fifty_carried_steps():
    """BASELINE configs[2]: GRU hidden 128, 512 streams, one frame per call, h carried over 50 calls."""
    cfg, m, sd = _model("gru")
    B, steps = 512, 50
    x = synth.features(B, steps, 80, seed=21)
    h = torch.zeros(2, B, 128)
    hd = h.to(DEV)
    worst = 0.0
    for t in range(steps):
        y, hd = m(x[:, t:t + 1].to(DEV), hd)
        y_ref, h = O.kws_forward(sd, cfg, x[:, t:t + 1], h)
        worst = max(worst, float((y.cpu() - y_ref).abs().max()))
        assert worst <= TOL_POST, (t, worst)
    assert float((hd.cpu() - h).abs().max()) <= TOL_POST
    print(f"gru B=512 T=1 x 50 steps: posterior max-abs {worst:.3e}")


@pytest.mark.parametrize("B,T", [(4, 40), (3, 13), (64, 40)])
def test_ds_tcn_ctc_vocabulary_2599(B, T):
    """The shipped CTC configuration (ds_tcn_ctc.yaml:31-42: hidden 256, output_dim 2599, identity activation):
    logits, softmax posteriors (forward_softmax, export_onnx.py:46-48) and the cache."""
    cfg, m, sd = _model("ds_tcn", activation="identity", output_dim=2599, input_dim=40)
    x = synth.features(B, T, 40, seed=8)
    cache = torch.randn(B, 256, 105, generator=torch.Generator().manual_seed(4))
    y, c = m(x.to(DEV), cache.to(DEV))
    p, _ = m.forward_softmax(x.to(DEV), cache.to(DEV))
    y_ref, c_ref = O.kws_forward(sd, cfg, x, cache)
    p_ref, _ = O.kws_forward(sd, cfg, x, cache, softmax=True)
    scale = max(1.0, float(y_ref.abs().max()))
    err, perr = float((y.cpu() - y_ref).abs().max()), float((p.cpu() - p_ref).abs().max())
    print(f"ds_tcn_ctc odim 2599 B={B} T={T}: logits max-abs {err:.3e} (|y|max {scale:.2f}), softmax max-abs {perr:.3e}")
    assert y.shape == (B, T, 2599)
    assert err <= TOL_POST * scale and perr <= TOL_POST
    assert float((c.cpu() - c_ref).abs().max()) <= TOL_POST * max(1.0, float(c_ref.abs().max()))


@pytest.mark.parametrize("B,T", [(1024, 40), (1, 300), (300, 1), (9, 57)])
@pytest.mark.parametrize("precision", ["auto", "fp32"])
def test_mdtc_small_large_tiles(B, T, precision):
    """hidden 32 tiles with more than 256 rows (B=1024 x T=40 packs 7 streams = 280 rows; one
    utterance of 300 frames) ran the time-parallel loops over the first 256 rows only."""
    cfg, m, sd = _model("mdtc_small", input_dim=40)
    m.precision = precision
    x = synth.features(B, T, 40, seed=31)
    cache = torch.randn(B, 32, 184, generator=torch.Generator().manual_seed(6))
    y, c = m(x.to(DEV), cache.to(DEV))
    y_ref, c_ref = O.kws_forward(sd, cfg, x, cache)
    err = float((y.cpu() - y_ref).abs().max())
    cerr = float((c.cpu() - c_ref).abs().max())
    assert err <= TOL_POST, (B, T, err)
    assert cerr <= TOL_POST * max(1.0, float(c_ref.abs().max())), (B, T, cerr)


def test_in_place_weight_edits_repack():
    """optimizer-style in-place edits and sub-module load_state_dict must not run stale packs."""
    cfg, m, sd = _model("tcn")
    x = synth.features(3, 24, 80, seed=2)
    y0, _ = m(x.to(DEV))
    with torch.no_grad():
        m.classifier.linear.bias.add_(0.75)
    sd2 = {k: v.clone().cpu() for k, v in m.state_dict().items()}
    y1, _ = m(x.to(DEV))
    assert float((y1 - y0).abs().max()) > 1e-3
    assert float((y1.cpu() - O.kws_forward(sd2, cfg, x, None)[0]).abs().max()) <= TOL_POST
    m.backbone.load_state_dict({k: torch.zeros_like(v) if k.endswith("cnn.0.weight") else v
                                for k, v in m.backbone.state_dict().items()})
    sd3 = {k: v.clone().cpu() for k, v in m.state_dict().items()}
    y2, _ = m(x.to(DEV))
    assert float((y2.cpu() - O.kws_forward(sd3, cfg, x, None)[0]).abs().max()) <= TOL_POST


def test_patch_reference_score_loop(tmp_path):
    """wekws/bin/score.py:109-137 restated around `wekws.model.kws_model` after patch_reference(): the reference's
    own import line resolves to this implementation, a reference-format checkpoint loads strictly, a zero-padded
    ragged batch is scored, and the '{:.6f}' score lines equal the oracle's to the text precision."""
    from wekws_b200 import patch_reference
    cfg, m0, sd = _model("mdtc", cmvn=True, output_dim=2)
    ckpt = tmp_path / "avg_3.pt"
    torch.save(sd, str(ckpt))
    patched_real = patch_reference()
    assert patched_real is False or os.path.isdir("/root/reference")
    from wekws.model.kws_model import init_model as ref_named_init        # score.py:30
    import wekws_b200.kws_model as ours
    assert ref_named_init is ours.init_model
    cmvn_file = synth.write_cmvn_json(80)
    try:
        configs = {"model": model_config("mdtc", cmvn_file=cmvn_file, output_dim=2)}
        model = ref_named_init(configs["model"])                          # score.py:108
    finally:
        os.unlink(cmvn_file)
    checkpoint = torch.load(str(ckpt), map_location="cpu")                # utils/checkpoint.py:23-30
    model.load_state_dict(checkpoint, strict=True)
    device = torch.device("cuda")                                         # score.py:110-113
    model = model.to(device)
    model.eval()
    lens = torch.tensor([98, 61, 98, 7, 33])
    pcm = synth.pcm_int16(5, 16000, seed=77)
    feats_cpu = torch.stack([O.fbank(pcm[b].float()) for b in range(5)])
    for b in range(5):
        feats_cpu[b, lens[b]:] = 0.0                                      # pad_sequence zero padding (processor.py)
    keys = ["utt%d" % i for i in range(5)]
    fout = io.StringIO()
    with torch.no_grad():                                                 # score.py:116-137
        feats = feats_cpu.to(device)
        feats_lengths = lens.to(device)
        logits, _ = model(feats)
        num_keywords = logits.shape[2]
        logits = logits.cpu()
        for i in range(len(keys)):
            score = logits[i][:feats_lengths[i]]
            for keyword_i in range(num_keywords):
                keyword_scores = score[:, keyword_i]
                score_frames = " ".join(["{:.6f}".format(x) for x in keyword_scores.tolist()])
                fout.write("{} {} {}\n".format(keys[i], "kw%d" % keyword_i, score_frames))
    y_ref, _ = O.kws_forward(sd, cfg, feats_cpu, None)
    lines = fout.getvalue().splitlines()
    assert len(lines) == 10
    for ln in lines:
        arr = ln.split()
        i, k = int(arr[0][3:]), int(arr[1][2:])
        got = np.array(list(map(float, arr[2:])))
        want = y_ref[i, :lens[i], k].numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= TOL_POST + 1e-6


# ------------------------------------------------------------------------------------------ FSMN
@pytest.mark.parametrize("case", ["fsmn", "fsmn_strided"])
def test_fsmn_streaming_matches_reference_golden(case):
    """Goldens made by the real reference FSMN (oracle/make_golden.py; they carry the weights): four chunks streamed
    with the 4-D cache carried (40, 17, 1, 9 frames) and the whole 67-frame utterance in one call."""
    from tests.cases import fsmn_config
    from tests.conftest import golden
    g = golden("model_" + case)
    cfg = fsmn_config(case)
    m = init_model(cfg).eval()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")}
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    assert m.backbone.padding == (cfg["backbone"]["left_order"] - 1) * cfg["backbone"]["left_stride"] + \
        cfg["backbone"]["right_order"] * cfg["backbone"]["right_stride"]
    cache = torch.zeros(0, 0, 0, 0)
    xs = []
    for i in range(4):
        x = torch.from_numpy(g[f"x{i}"])
        xs.append(x)
        y, cache = m(x.to(DEV), cache)
        assert y.shape == g[f"y{i}"].shape and cache.shape == g[f"c{i}"].shape
        assert np.abs(y.cpu().numpy() - g[f"y{i}"]).max() <= TOL_POST * max(1.0, float(np.abs(g[f"y{i}"]).max())), (case, i)
        assert np.abs(cache.cpu().numpy() - g[f"c{i}"]).max() <= TOL_POST * max(1.0, float(np.abs(g[f"c{i}"]).max()))
    yf, _ = m(torch.cat(xs, dim=1).to(DEV))
    assert np.abs(yf.cpu().numpy() - g["y_full"]).max() <= TOL_POST * max(1.0, float(np.abs(g["y_full"]).max()))


@pytest.mark.parametrize("B,T", [(3, 50), (64, 1), (5, 130), (200, 7)])
def test_fsmn_shipped_size_matches_oracle(B, T):
    """fsmn_ctc.yaml shape (400 -> 140 -> 250, 4 x (128-dim memory, orders 10 / 2), -> 140 -> 2599) with random weights
    and a random cache, logits and softmax posteriors against the oracle (pinned to the reference by the goldens);
    in-place cache update (out_cache aliasing in_cache is what a streaming caller does)."""
    cfg = model_config("fsmn", input_dim=400, output_dim=2599)
    torch.manual_seed(777)
    m = synth.randomize_(init_model(cfg), seed=777).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    x = synth.features(B, T, 400, seed=12)
    cache = torch.randn(B, 128, 11, 4, generator=torch.Generator().manual_seed(9))
    y, c = m(x.to(DEV), cache.to(DEV))
    p, _ = m.forward_softmax(x.to(DEV), cache.to(DEV))
    y_ref, c_ref = O.kws_forward(sd, cfg, x, cache)
    p_ref, _ = O.kws_forward(sd, cfg, x, cache, softmax=True)
    scale = max(1.0, float(y_ref.abs().max()))
    err, perr = float((y.cpu() - y_ref).abs().max()), float((p.cpu() - p_ref).abs().max())
    print(f"fsmn shipped size B={B} T={T}: logits max-abs {err:.3e} (|y|max {scale:.2f}), softmax max-abs {perr:.3e}")
    assert y.shape == (B, T, 2599) and c.shape == (B, 128, 11, 4)
    assert err <= TOL_POST * scale and perr <= TOL_POST
    assert float((c.cpu() - c_ref).abs().max()) <= TOL_POST * max(1.0, float(c_ref.abs().max()))
    # two half chunks == one chunk (streaming), empty cache == zero cache
    if T >= 2:
        h = T // 2
        y1, c1 = m(x[:, :h].to(DEV), cache.to(DEV))
        y2, c2 = m(x[:, h:].to(DEV), c1)
        assert float((torch.cat((y1, y2), 1) - y).abs().max()) <= TOL_POST * scale
        assert float((c2 - c).abs().max()) <= TOL_POST * max(1.0, float(c_ref.abs().max()))
    y0, _ = m(x.to(DEV))
    yz, _ = m(x.to(DEV), torch.zeros_like(cache).to(DEV))
    assert torch.equal(y0, yz)


def test_context_expansion_and_frame_skip_bit_exact():
    """Device transform == reference processor (golden) and oracle, bit for bit; ragged batch with zero padding."""
    from tests.conftest import golden
    from wekws_b200 import context_expansion
    g = golden("context")
    for i, (T, D, left, right, skip) in enumerate(g["cases"].tolist()):
        y, n = context_expansion(torch.from_numpy(g[f"x{i}"]).unsqueeze(0).to(DEV), left, right, skip)
        assert torch.equal(y[0].cpu(), torch.from_numpy(g[f"y{i}"])) and int(n[0]) == g[f"y{i}"].shape[0], i
    x = synth.features(5, 50, 80, seed=3)
    lens = torch.tensor([50, 49, 7, 2, 0])
    y, n = context_expansion(x.to(DEV), 2, 2, 3, lengths=lens)
    assert y.shape == (5, 16, 400)
    for b in range(5):
        want = O.frame_skip(O.context_expansion(x[b, :lens[b]], 2, 2), 3) if lens[b] > 2 else torch.zeros(0, 400)
        assert int(n[b]) == want.shape[0]
        assert torch.equal(y[b, :want.shape[0]].cpu(), want) and float(y[b, want.shape[0]:].abs().sum()) == 0.0


# ------------------------------------------------------------------------------- CTC prefix beam search
KEYWORDS_TOKEN = {"hi_xiaowen": {"token_id": [5, 9, 17, 23]}, "nihao_wenwen": {"token_id": [31, 7, 23, 23]}}


def test_ctc_prefix_beam_search_bit_exact_with_reference_golden(tmp_path):
    """Device decoder == the hypotheses returned by the reference's own loss.py ctc_prefix_beam_search (golden) and the
    oracle: beam order, prefixes, pb + pnb as doubles (exact), node frames / probabilities; ragged batch; chunked
    decoding with carried state == whole utterance; keyword look-up and the score_ctc.py score lines."""
    import io
    from tests.test_oracle_pinned import _ctc_golden_cases
    from wekws_b200 import ctc_keyword_hits, ctc_prefix_beam_search, ctc_state, write_ctc_scores
    cases = list(_ctc_golden_cases())
    for use_set in (True, False):
        sel = [c for c in cases if (c[2] is not None) == use_set]
        if not sel:
            continue
        tokenset = sel[0][2]
        Tm = max(c[1].size(0) for c in sel)
        V = sel[0][1].size(1)
        probs = torch.zeros(len(sel), Tm, V)
        lens = torch.tensor([c[1].size(0) for c in sel], dtype=torch.int32)
        for k, c in enumerate(sel):
            probs[k, :c[1].size(0)] = c[1]
        res = ctc_prefix_beam_search(probs.to(DEV), lens, tokenset)
        assert int(res.overflow.sum()) == 0
        got = res.to_python()
        for k, c in enumerate(sel):
            assert got[k] == c[3], (use_set, c[0])
        # chunked with carried hypotheses (the streaming caller, stream_kws_ctc.py:482-501)
        st = ctc_state(len(sel), DEV)
        for t0 in range(0, Tm, 23):
            part = probs[:, t0:t0 + 23]
            r2 = ctc_prefix_beam_search(part.to(DEV), (lens - t0).clamp(0, part.size(1)), tokenset, state=st,
                                        reset_state=(t0 == 0), frame_offset=t0)
        assert r2.to_python() == got
        # keyword look-up + score file
        hits = ctc_keyword_hits(res, KEYWORDS_TOKEN)
        want = [O.ctc_keyword_hit(c[3], KEYWORDS_TOKEN) for c in sel]
        assert hits == want
        a, b = io.StringIO(), io.StringIO()
        write_ctc_scores(a, ["utt%d" % k for k in range(len(sel))], hits)
        for k, (word, sc, _, _) in enumerate(want):                   # score_ctc.py:217-226
            b.write('{} detected {} {:.3f}\n'.format("utt%d" % k, word, sc) if word is not None else '{} rejected\n'.format("utt%d" % k))
        assert a.getvalue() == b.getvalue()


def test_ctc_decode_large_vocabulary_matches_oracle():
    """Vocabulary 2599 (the shipped CTC token list size), 64 utterances: device decoder == oracle restatement."""
    from wekws_b200 import ctc_keyword_hits, ctc_prefix_beam_search
    g = torch.Generator().manual_seed(17)
    B, T, V = 64, 80, 2599
    kw = {"kw_a": {"token_id": [100, 2000, 57]}, "kw_b": {"token_id": [2598, 100]}}
    tokenset = {0, 100, 2000, 57, 2598}
    logits = torch.randn(B, T, V, generator=g) * 0.3
    dom = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        t = int(torch.randint(0, 10, (1,), generator=g))
        for tok in ([100, 2000, 57] if b % 3 else [2598, 100, 2000, 57, 57]):
            for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
                if t < T:
                    dom[b, t] = tok
                    t += 1
            t += int(torch.randint(0, 3, (1,), generator=g))
    logits.scatter_add_(2, dom.unsqueeze(2), torch.full((B, T, 1), 12.0))
    probs = logits.softmax(2)
    res = ctc_prefix_beam_search(probs.to(DEV), None, tokenset)
    got = res.to_python()
    hits = ctc_keyword_hits(res, kw)
    nhit = 0
    for b in range(B):
        want = O.hyps_of(O.ctc_prefix_beam_search(probs[b], tokenset))
        assert got[b] == want, b
        assert hits[b] == O.ctc_keyword_hit(want, kw)
        nhit += hits[b][0] is not None
    assert nhit >= B // 2


@pytest.mark.parametrize("B,T,idim", [(70, 40, 40), (1, 1, 40), (512, 3, 80), (1300, 2, 40), (9600, 1, 80), (33, 7, 72)])
def test_gru_tensor_core_path(B, T, idim):
    """The weight-streaming tcgen05 GRU (gru_tc.cu: transposed GEMMs, units on the TMEM lanes) against the oracle and
    against the FP32 kernel: ragged tiles (B not a multiple of 64), more tiles than SMs, one and two feature K slabs,
    random h0, CMVN, two outputs."""
    cfg, m, sd = _model("gru", cmvn=True, input_dim=idim, output_dim=2)
    x = synth.features(B, T, idim, seed=5, cmvn_like=True)
    h0 = torch.randn(2, B, 128, generator=torch.Generator().manual_seed(8)) * 0.5
    try:
        m.precision = "fp32"
        y32, h32 = m(x.to(DEV), h0.to(DEV))
        assert not m.uses_tensor_cores(T)
        m.precision = "tensor"                      # the tcgen05 kernel whatever the batch ("auto" picks by B and T)
        ytc, htc = m(x.to(DEV), h0.to(DEV))
        assert m.uses_tensor_cores(T, B)
        m.precision = "auto"
        assert m.uses_tensor_cores(T, B) == (B >= (640 if T == 1 else 400 if T < 8 else 256))
        ya, ha = m(x.to(DEV), h0.to(DEV))
        assert torch.equal(ya, ytc if m.uses_tensor_cores(T, B) else y32)
        m.precision = "tensor"
    finally:
        pass
    y_ref, h_ref = O.kws_forward(sd, cfg, x, h0)
    e32, etc = float((y32.cpu() - y_ref).abs().max()), float((ytc.cpu() - y_ref).abs().max())
    print(f"gru B={B} T={T} idim={idim}: posterior max-abs fp32 kernel {e32:.2e}, tensor-core kernel {etc:.2e}")
    assert e32 <= TOL_POST and etc <= TOL_POST
    assert float((htc.cpu() - h_ref).abs().max()) <= TOL_POST and float((h32.cpu() - h_ref).abs().max()) <= TOL_POST
    # streaming: T calls of one frame == one call of T frames (state carried)
    hd, ys = h0.to(DEV), []
    for t in range(T):
        yt, hd = m(x[:, t:t + 1].to(DEV), hd)
        ys.append(yt)
    m.precision = "auto"
    assert float((torch.cat(ys, 1) - ytc).abs().max()) <= 2e-5 and float((hd - htc).abs().max()) <= 2e-5


def test_pipeline_native_call_matches_two_step_chain_and_oracle():
    """Pipeline(frontend, model)(pcm) == model(frontend(pcm)) bit for bit (same kernels, L2-pinned features), for Fbank
    and MFCC front-ends, with a carried cache, and against the oracle at the posterior gate."""
    from wekws_b200 import Pipeline
    cfg, m, sd = _model("mdtc", cmvn=True)
    pcm = synth.pcm_int16(33, 16000, seed=4)
    for fe in (Fbank(80), Mfcc(80, 80)):
        pipe = Pipeline(fe, m)
        y1, c1 = pipe(pcm.to(DEV))
        y2, c2 = m(fe(pcm.to(DEV)))
        assert torch.equal(y1, y2) and torch.equal(c1, c2)
        y3, c3 = pipe(pcm.to(DEV), c1)
        y4, c4 = m(fe(pcm.to(DEV)), c2)
        assert torch.equal(y3, y4) and torch.equal(c3, c4)
    ref_f = torch.stack([O.mfcc(pcm[b].float(), 80, 80) for b in range(pcm.shape[0])])
    y_ref, _ = O.kws_forward(sd, cfg, ref_f, None)
    assert float((y1.cpu() - y_ref).abs().max()) <= TOL_POST
    with pytest.raises(ValueError):
        Pipeline(Fbank(40), m)
