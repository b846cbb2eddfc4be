"""The contraction primitive under every Linear of the MLP stage (sln_contract): tcgen05 3xTF32 tiles and FP32 SIMT tiles
against an fp64 matmul, in all four operand orientations, plain and split-K accumulate modes."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
_lib = importlib.import_module("sln_b200._lib")
DEV = "cuda:0"


def _run(M, N, K, a_rc, b_rc, accumulate, engine, seed=0):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    A[:, K // 2:] *= 37.0                       # mixed magnitudes: exercises the hi/lo split
    want = A.double() @ B.double().t()
    Ad = (A if a_rc else A.t().contiguous()).to(DEV)
    Bd = (B if b_rc else B.t().contiguous()).to(DEV)
    C0 = torch.randn(M, N, generator=g)
    C = C0.to(DEV) if accumulate else torch.full((M, N), float("nan"), device=DEV)
    _lib.check(lib.sln_contract(Ad.data_ptr(), Ad.stride(0), int(a_rc), Bd.data_ptr(), Bd.stride(0), int(b_rc), C.data_ptr(), N, M, N, K,
                                int(accumulate), engine, _lib.cur_stream(torch.device(DEV))), "contract")
    torch.cuda.synchronize()
    if accumulate:
        want = want + C0.double()
    err = (C.cpu().double() - want).abs().max().item() / want.abs().max().item()
    return err


SHAPES = [(128, 64, 32), (128, 256, 64), (256, 128, 96), (3968, 640, 256), (2048, 128, 256), (300, 100, 76), (132, 36, 40), (640, 384, 3968)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_rc,b_rc", [(True, True), (True, False), (False, False), (False, True)])
def test_tcgen05_3xtf32_matches_fp64(M, N, K, a_rc, b_rc):
    err = _run(M, N, K, a_rc, b_rc, False, 1)
    assert err < 3e-6, err          # fp32-level accuracy: single-pass TF32 would be ~5e-4


@pytest.mark.parametrize("M,N,K", [(640, 256, 3968), (256, 384, 3968), (128, 256, 2048), (200, 72, 1000), (2048, 128, 128), (64, 32, 32)])
def test_tcgen05_split_k_accumulate(M, N, K):
    assert _run(M, N, K, False, False, True, 1) < 3e-6


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (3968, 640, 256), (300, 100, 77), (5, 6, 7)])
@pytest.mark.parametrize("a_rc,b_rc", [(True, True), (True, False), (False, False)])
def test_simt_fp32_matches_fp64(M, N, K, a_rc, b_rc):
    assert _run(M, N, K, a_rc, b_rc, False, 0) < 3e-6
