"""tcgen05 linear kernel (TMA-fed, error-free bf16x3 slices; csrc/tc_gemm.cuh) vs fp64 numpy, and vs the FP32
SIMT kernel inside the engine."""
import ctypes

import numpy as np
import pytest

from helpers import model_weights, species_of

pytestmark = pytest.mark.gpu


def _dense(rows, K, N, use_tc, seed=0):
    import torch
    from sevenn_b200.engine import check, load_library
    lib = load_library()
    rng = np.random.RandomState(seed)
    A = rng.normal(size=(rows, K)).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    a, w = torch.tensor(A, device='cuda'), torch.tensor(W, device='cuda')
    c = torch.full((rows, N), float('nan'), device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.s7b_dense_linear(a.data_ptr(), w.data_ptr(), c.data_ptr(), rows, K, N, use_tc, st))
    torch.cuda.synchronize()
    return c.cpu().numpy(), A.astype(np.float64) @ W.astype(np.float64)


@pytest.mark.parametrize('rows,K,N', [(128, 32, 32), (300, 64, 64), (1000, 224, 224), (257, 416, 64),
                                      (640, 352, 32), (513, 128, 384), (128, 32, 16), (4096, 384, 256)])
def test_tc_linear_matches_fp64(rows, K, N):
    got, ref = _dense(rows, K, N, 1, seed=rows + K)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max()
    # fp32-level: a plain fp32 dot product of length K has ~1e-7 * sqrt(K) relative error
    assert err < 4e-7 * np.sqrt(K) * max(1.0, np.abs(ref).max()), err
    simt, _ = _dense(rows, K, N, 0, seed=rows + K) if K % 4 == 0 else (got, None)
    assert np.abs(got - simt).max() < 6e-7 * np.sqrt(K) * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('rows,K,N', [(300, 64, 64), (1000, 224, 224), (77, 8, 64)])
def test_simt_linear_matches_fp64(rows, K, N):
    got, ref = _dense(rows, K, N, 0, seed=1)
    assert np.abs(got - ref).max() < 3e-6 * np.sqrt(K) * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('rows,K,N', [(1000, 224, 224), (5000, 384, 64)])
def test_tc_linear_has_no_accumulation_bias(rows, K, N):
    """The tensor core's truncating accumulation biased the round-1 3xTF32 kernel (~ -3e-7 relative); the
    fixed-point slices make the first-order accumulator exact, so the signed error averages to zero."""
    got, ref = _dense(rows, K, N, 1, seed=11)
    scale = np.abs(ref).mean()
    signed = (got - ref) * np.sign(ref) / scale
    assert abs(signed.mean()) < 4.0 * signed.std() / np.sqrt(signed.size) + 1e-10
    assert np.sqrt(((got - ref) ** 2).mean()) / scale < 3e-6


def test_tc_linear_swizzled_and_plain_tma_tiles_agree():
    from sevenn_b200.engine import set_option
    try:
        set_option('tc_swizzle', 0)
        plain, ref = _dense(777, 224, 112, 1, seed=3)
        set_option('tc_swizzle', 1)
        swz, _ = _dense(777, 224, 112, 1, seed=3)
    finally:
        set_option('tc_swizzle', 1)
    assert np.array_equal(plain, swz)
    assert np.abs(swz - ref).max() < 4e-7 * np.sqrt(224) * max(1.0, np.abs(ref).max())


def test_engine_tc_and_simt_linears_agree():
    import torch
    from sevenn_b200.engine import B200Engine, set_option
    from sevenn_b200.neighbors import build_graph, diamond_si
    meta, arrays = model_weights('sevennet_0')
    pos, cell, z = diamond_si(3, 3, 3)
    ei, ev = build_graph(pos, cell, True, 5.0)
    e = B200Engine(meta, arrays)
    e.set_graph(species_of(meta, z), ei, ev)
    out = {}
    try:
        for tc in (0, 1):
            set_option('tc_gemm', tc)
            e.compute()
            torch.cuda.synchronize()
            r = e.results()
            out[tc] = (float(r['energy'].cpu()[0]), r['forces'].cpu().numpy())
    finally:
        set_option('tc_gemm', 1)
    # 216 atoms: no systematic per-atom energy shift between the two GEMM paths (round 1: -8.7e-6 eV/atom)
    assert abs(out[0][0] - out[1][0]) < 216 * 3e-7, (out[0][0], out[1][0])
    assert np.allclose(out[0][1], out[1][1], atol=2e-5)
