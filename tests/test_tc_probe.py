"""GPU probe of the tcgen05 / TMEM / UMMA-descriptor building blocks (csrc/tc_common.cuh):
a single 128x64xK bf16x3 GEMM through shared-memory operands in the SW128 K-major layout."""
import ctypes as C
import os

import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
LIB = os.path.join(ROOT, "tests", "native", "libtc_probe.so")


@pytest.mark.parametrize("K", [64, 80, 48, 40])
def test_umma_bf16x3_gemm_matches_fp32(K):
    assert os.path.exists(LIB), "tests/native/libtc_probe.so not built (run __graft_entry__.build())"
    lib = C.CDLL(LIB)
    lib.tc_probe_run.restype = C.c_int
    lib.tc_probe_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    g = torch.Generator().manual_seed(K)
    A = (torch.randn(128, K, generator=g) * 3).cuda()
    W = (torch.randn(64, K, generator=g) * 0.3).cuda()
    D = torch.full((128, 64), float("nan"), device="cuda")
    assert lib.tc_probe_run(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(D.data_ptr()), K) == 0
    ref = (A.double() @ W.double().T)
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ W.abs().double().T).max().item()
    assert err <= 4e-5 * scale, (err, scale)          # bf16x3: ~2^-16 relative to sum |a||w|
    # and clearly better than a single bf16 pass (2^-8): proves all three passes accumulate
    assert err <= 1e-3 * scale


def test_umma_a_operand_from_tmem():
    """tcgen05.mma with the A operand in TMEM (written with tcgen05.st): rows = lanes, two bf16 per column."""
    lib = C.CDLL(LIB)
    lib.tc_probe_ts_run.restype = C.c_int
    lib.tc_probe_ts_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    g = torch.Generator().manual_seed(7)
    A = (torch.randn(128, 64, generator=g) * 3).cuda()
    W = (torch.randn(64, 64, generator=g) * 0.3).cuda()
    D = torch.full((128, 64), float("nan"), device="cuda")
    assert lib.tc_probe_ts_run(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(D.data_ptr())) == 0
    ref = (A.double() @ W.double().T)
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ W.abs().double().T).max().item()
    assert err <= 4e-5 * scale, (err, scale)
