"""CPU tests of the ONNX-contract exporter (wekws_b200/export_onnx.py, the role of wekws/bin/export_onnx.py): the file
is decoded again with tests/onnx_mini.py and run the way runtime/core/kws/keyword_spotting.cc:56-95 runs it
(batch 1, chunk by chunk, r_cache fed back as cache) against the oracle and the reference-made goldens."""
import numpy as np
import pytest
import torch

from oracle import kws_oracle as O
from tests import onnx_mini
from tests.cases import build_model, fsmn_config
from tests.conftest import golden
from wekws_b200 import export_onnx, init_model, synth

ATOL = 1e-6          # export_onnx.py:88-91 checks torch vs ORT with allclose(atol=1e-6) (rtol 1e-5)


def _close(a, b, scale=1.0):
    return (a - b).abs().max().item() <= ATOL * 10 * max(1.0, scale) + 1e-5 * b.abs().max().item()


@pytest.mark.parametrize("case", ["mdtc", "mdtc_cmvn_logits", "mdtc_small", "tcn", "ds_tcn", "ds_tcn_ctc"])
def test_onnx_file_keeps_the_runtime_contract_and_the_oracle_numbers(case, tmp_path):
    cfg, model, _ = build_model(case, init_model)
    softmax = case == "ds_tcn_ctc"                    # CTC models export forward_softmax (export_onnx.py:46-48)
    path = str(tmp_path / (case + ".onnx"))
    meta = export_onnx(model, path, softmax=softmax)
    m = onnx_mini.load(path)
    # the contract keyword_spotting.cc:28-46 binds to
    assert m["opset"] == 13 and m["ir_version"] == 7
    assert [n for n, _ in m["inputs"]] == ["input", "cache"] and [n for n, _ in m["outputs"]] == ["output", "r_cache"]
    P = model.backbone.padding
    assert m["inputs"][0][1] == [1, "T", cfg["input_dim"]] and m["inputs"][1][1] == [1, model.hdim, P]
    assert m["outputs"][0][1] == [1, "T", cfg["output_dim"]] and m["outputs"][1][1] == [1, model.hdim, P]
    assert m["metadata"] == {"cache_dim": str(model.hdim), "cache_len": str(P)} == \
        {k: meta[k] for k in ("cache_dim", "cache_len")}
    assert not any(n["op"] == "BatchNormalization" for n in m["nodes"])          # folded, as constant folding does
    # streaming run: zero cache at Reset() (keyword_spotting.cc:47-54), r_cache fed back
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    has_cmvn = model.global_cmvn is not None
    cache = torch.zeros(1, model.hdim, P)
    ref_cache = None
    for i, T in enumerate((100, 40, 17, 1)):
        x = synth.features(1, T, cfg["input_dim"], seed=10 + i, cmvn_like=has_cmvn)
        y, cache = onnx_mini.run(m, {"input": x, "cache": cache})
        y_ref, ref_cache = O.kws_forward(sd, cfg, x, ref_cache, softmax=softmax)
        assert y.shape == y_ref.shape and cache.shape == ref_cache.shape
        assert _close(y, y_ref), (case, i, (y - y_ref).abs().max().item())
        assert _close(cache, ref_cache, ref_cache.abs().max().item()), (case, i)


@pytest.mark.parametrize("case", ["fsmn", "fsmn_strided"])
def test_onnx_fsmn_matches_the_reference_goldens(case, tmp_path):
    """4-D cache (1, proj, cache_len, layers) (export_onnx.py:59-60); numbers from the real reference FSMN."""
    g = golden("model_" + case)
    cfg = fsmn_config(case)
    model = init_model(cfg).eval()
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")}, strict=True)
    path = str(tmp_path / "fsmn.onnx")
    export_onnx(model, path)
    m = onnx_mini.load(path)
    bb = model.backbone
    shape = [1, bb.proj_dim, bb.cache_len, bb.fsmn_layers]
    assert m["inputs"][1] == ("cache", shape) and m["outputs"][1] == ("r_cache", shape)
    for b in range(g["x0"].shape[0]):                 # the runtime is batch 1: run every golden stream on its own
        cache = torch.zeros(shape)
        for i in range(4):
            x = torch.from_numpy(g[f"x{i}"][b:b + 1])
            y, cache = onnx_mini.run(m, {"input": x, "cache": cache})
            assert np.abs(y.numpy() - g[f"y{i}"][b:b + 1]).max() <= 1e-5 * max(1.0, np.abs(g[f"y{i}"]).max()), (case, b, i)
            assert np.abs(cache.numpy() - g[f"c{i}"][b:b + 1]).max() <= 1e-5


def test_onnx_export_refuses_gru_like_the_reference(tmp_path):
    _, model, _ = build_model("gru", init_model)
    with pytest.raises(NotImplementedError):
        export_onnx(model, str(tmp_path / "gru.onnx"))


def test_protobuf_varints_and_negative_slice_bounds_roundtrip(tmp_path):
    """Wire-format details a real ONNX parser is strict about: negative int64 as 10-byte varints, dims, raw_data."""
    from wekws_b200.export_onnx import _Graph, _attr_ints, _f_bytes, _tensor, _varint
    assert _varint(0) == b"\x00" and _varint(300) == b"\xac\x02" and len(_varint(-1)) == 10
    name, t = onnx_mini._tensor(_tensor("w", np.arange(6, dtype=np.float32).reshape(2, 3)))
    assert name == "w" and t.shape == (2, 3) and t[1, 2] == 5
    name, t = onnx_mini._tensor(_tensor("i", np.array([-7, np.iinfo(np.int64).max], dtype=np.int64)))
    assert t.tolist() == [-7, np.iinfo(np.int64).max]
    a = onnx_mini.fields(_attr_ints("pads", [0, -3]))
    assert [onnx_mini._signed(v) for f, _, v in a if f == 8] == [0, -3]
