"""Edge cases of the hot path on the GPU against the oracle: smallest grid, clamped skewness values, a single
Q2 point, backward and below-charm Q2 schedules, repeated calls, empty / strided application inputs and the
call-order contract of the reference (matrices.py:40-41,59-68,142-145; matevo.f90:70,688-689)."""
import ctypes as C

import numpy as np
import pytest
import torch

from test_gpu_parity import close, _raw_gpu, _qags_gpu

pytestmark = pytest.mark.gpu

NAMES = (("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}), ("matrix_ASG", {}))


def _same_kernels(tk, tko, nfls=(3, 4, 5)):
    for nfl in nfls:
        for which, name in enumerate(tko.RAW_NAMES):
            got = _raw_gpu(tk, which, nfl)
            ref = tko.raw_kernel(name, nfl)
            assert np.array_equal(got, ref, equal_nan=True), "%s nfl=%d not bit-identical" % (name, nfl)


@pytest.mark.parametrize("grid_type", [1, 2])
def test_smallest_grid_nx6(tk, tko, grid_type):
    """nx = 6 is the smallest grid the reference accepts (matrices.py:40-41): every stencil is the whole grid."""
    xis = np.array([0.25])
    tk.matrices.initialize_kernels(6, xis, grid_type=grid_type)
    tko.initialize_kernels(6, xis, grid_type, mode=0)
    _same_kernels(tk, tko)
    Q2 = np.array([4.0, 8.0])
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    tko.initialize_evolution_matrices(Q2, nlo=True)
    for name, kw in NAMES:
        close(getattr(tk.matrices, name)(**kw).cpu().numpy(), getattr(tko, name)(**kw))


def test_odd_or_tiny_nx_rejected(tk):
    with pytest.raises(AssertionError):
        tk.matrices.initialize_kernels(7, np.array([0.5]))
    with pytest.raises(AssertionError):
        tk.matrices.initialize_kernels(4, np.array([0.5]))


@pytest.mark.parametrize("grid_type", [1, 2])
def test_xi_clamp_on_first_and_last_element_only(tk, tko, grid_type):
    """initialize_2D clamps xi(1) >= 1e-6 and xi(nxi) <= 0.999999 and nothing else (matevo.f90:688-689)."""
    xis = np.array([1e-7, 0.3, 0.9999999])
    tk.matrices.initialize_kernels(10, xis, grid_type=grid_type)
    tko.initialize_kernels(10, xis, grid_type, mode=1)
    _same_kernels(tk, tko, nfls=(4,))
    # the clamped values are what was assembled: the same as asking for them directly
    a = _raw_gpu(tk, 0, 4).copy()
    tk.matrices.initialize_kernels(10, np.array([1e-6, 0.3, 0.999999]), grid_type=grid_type)
    assert np.array_equal(a, _raw_gpu(tk, 0, 4), equal_nan=True)
    # QAGS paths at the clamped end points
    _, last, ifail = _qags_gpu(tk)
    nodes = tko.grid_nodes(10, 1e-6, grid_type)
    for i in range(10):
        for c in range(7):
            _, l, e = tko.cst_integral(c, nodes[i], 1e-6)
            assert list(last[0, i, c]) == l and list(ifail[0, i, c]) == e


def test_single_q2_point_rejected_and_first_slab_is_identity(tk):
    """matrices.py:68: at least two Q2 points; M(xi, Q2_1) = 1 (matevo.f90:505-510)."""
    tk.matrices.initialize_kernels(8, np.array([0.2, 0.4]))
    with pytest.raises(AssertionError):
        tk.matrices.initialize_evolution_matrices(np.array([4.0]), nlo=True)
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 4.0]), nlo=True)
    for name, kw in NAMES:
        M = getattr(tk.matrices, name)(**kw)
        m = M.shape[0]
        assert tuple(M.shape) == (m, m, 2, 2)
        eye = torch.eye(m, dtype=torch.float64, device=M.device)[:, :, None].expand(m, m, 2)
        assert torch.equal(M[:, :, :, 0], eye)
        # a zero-length step: alpha_s is finite, log(Q2_2/Q2_1) = 0, so the RK4 increment vanishes exactly
        assert torch.equal(M[:, :, :, 1], eye)


@pytest.mark.parametrize("Q2", [
    [10.0, 4.0, 2.0],                       # backward evolution (decreasing Q2), matevo.f90:512-516
    [1.2, 1.5, 1.6641000000000001, 3.0],    # starts below mc2 (nfl = 3) and lands exactly on the threshold
    [4.0, 4.0, 17.4724, 17.4724, 30.0],     # repeated points and the exact mb2 value
])
def test_q2_schedules(tk, tko, Q2):
    nx, xis = 12, np.array([0.15, 0.55])
    Q2 = np.array(Q2)
    tk.matrices.initialize_kernels(nx, xis, grid_type=2)
    tko.initialize_kernels(nx, xis, 2, mode=1)
    for nlo in (False, True):
        tk.matrices.initialize_evolution_matrices(Q2, nlo=nlo)
        tko.initialize_evolution_matrices(Q2, nlo=nlo)
        for name, kw in NAMES:
            close(getattr(tk.matrices, name)(**kw).cpu().numpy(), getattr(tko, name)(**kw))


def test_kernels_persist_across_make_matrices_and_are_rebuilt_by_make_kernels(tk, tko):
    """Call-order contract (matrices.py:59-66, matevo.f90:70): matrices can be rebuilt on the same kernels;
    new kernels drop the matrices."""
    nx = 10
    tk.matrices.initialize_kernels(nx, np.array([0.3]))
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 9.0]), nlo=False)
    a = tk.matrices.matrix_VSG().clone()
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 6.0, 9.0]), nlo=False)
    b = tk.matrices.matrix_VSG()
    assert tuple(b.shape) == (2 * nx, 2 * nx, 1, 3)
    # one RK4 step 4 -> 9 is not two steps 4 -> 6 -> 9, but both are the same evolution to O(step^5)
    assert torch.allclose(a[..., 1], b[..., 2], rtol=0, atol=1e-3 * float(a[..., 1].abs().max()))
    tk.matrices.initialize_kernels(nx, np.array([0.3, 0.5]))
    L = tk._lib.lib()
    assert L.tkb_get_nq2() == 0
    p, shape, strides = C.c_void_p(), (C.c_int64 * 4)(), (C.c_int64 * 4)()
    assert L.tkb_evomatrix(1, 0, C.byref(p), shape, strides) != 0
    assert b"make_matrices" in L.tkb_last_error() or b"matrices" in L.tkb_last_error()


def test_getter_copies_are_private(tk):
    """f2py intent(out) semantics: the caller may scribble on what a getter returned (matrices.py:146-147)."""
    tk.matrices.initialize_kernels(8, np.array([0.4]))
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 5.0]), nlo=False)
    M = tk.matrices.matrix_VNS()
    keep = M.clone()
    M.zero_()
    assert torch.equal(tk.matrices.matrix_VNS(), keep)


def test_apply_empty_and_strided_inputs(tk):
    nx, nxi, nq = 12, 3, 4
    tk.matrices.initialize_kernels(nx, np.array([0.1, 0.3, 0.6]))
    tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 40.0, nq), nlo=True)
    M = tk.matrices.matrix_VSG()
    m = 2 * nx
    gen = torch.Generator(device="cuda").manual_seed(7)
    # a non-contiguous sample block (every second sample of a larger array) and a transposed one
    big = torch.rand((m, nxi, 10), generator=gen, dtype=torch.float64, device="cuda")
    g = big[:, :, ::2]
    out = tk.matrices.apply_matrix(M, g)
    ref = torch.einsum("ijkl,jks->iksl", M, g)
    assert tuple(out.shape) == (m, nxi, 5, nq)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12 * float(ref.abs().max()))
    gt = torch.rand((nxi, m), generator=gen, dtype=torch.float64, device="cuda").t()        # (m, nxi), strided
    out = tk.matrices.apply_matrix(M, gt)
    ref = torch.einsum("ijkl,jk->ikl", M, gt)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12 * float(ref.abs().max()))
    # no samples: an empty result of the right shape, no launch failure
    empty = torch.empty((m, nxi, 0), dtype=torch.float64, device="cuda")
    out = tk.matrices.apply_matrix(M, empty)
    assert tuple(out.shape) == (m, nxi, 0, nq)
    # NaN / Inf in the input propagate like in einsum (0 * Inf = NaN included)
    gn = torch.rand((m, nxi, 2), generator=gen, dtype=torch.float64, device="cuda")
    gn[3, 1, 0] = float("nan")
    gn[5, 2, 1] = float("inf")
    out = tk.matrices.apply_matrix(M, gn)
    ref = torch.einsum("ijkl,jks->iksl", M, gn)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    fin = torch.isfinite(ref)
    assert torch.allclose(out[fin], ref[fin], rtol=1e-12, atol=1e-12 * float(ref[fin].abs().max()))


def test_apply_more_than_65535_matrices_with_shared_right_hand_side(tk):
    """nxi * nQ2 > 65535 with nQ2 = 16 (e.g. the gathered C5 non-singlet set, 4096 xi x 16 Q2): the batch goes to the
    GEMM in gridDim.z chunks, which must start at a multiple of nQ2 because the nQ2 matrices of a xi share one
    right-hand side."""
    m, nxi, nq, S = 40, 4097, 16, 3
    gen = torch.Generator(device="cuda").manual_seed(11)
    M = torch.rand((nxi, nq, m, m), generator=gen, dtype=torch.float64, device="cuda") - 0.5
    g = torch.rand((m, nxi, S), generator=gen, dtype=torch.float64, device="cuda")
    out = torch.empty((nxi, nq, m, S), dtype=torch.float64, device="cuda")
    L = tk._lib.lib()
    tk._lib.check(L.tkb_apply(m, nxi, nq, S, tk._lib.ptr(M), tk._lib.ptr(g), nxi * S, tk._lib.ptr(out), tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.einsum("kqij,jks->kqis", M, g)
    assert float((out - ref).abs().max()) < 1e-12 * float(ref.abs().max())


def test_many_xi_few_rows_queue_drains(tk, tko):
    """More QAGS instances than slots and fewer than one full round of them at the end: the persistent
    blocks must drain the queue and terminate (both groups of the pipeline, all constants)."""
    nx = 6
    xis = np.geomspace(1e-3, 0.8, 37)
    tk.matrices.initialize_kernels(nx, xis, grid_type=2)
    res, last, ifail = _qags_gpu(tk)
    assert (last >= 1).all() and (last <= 500).all()
    for k in (0, 17, 36):
        nodes = tko.grid_nodes(nx, xis[k], 2)
        for i in range(nx):
            for c in range(7):
                v, l, e = tko.cst_integral(c, nodes[i], xis[k])
                assert list(last[k, i, c]) == l and list(ifail[k, i, c]) == e
                got = res[k, i, c]
                tot = ((0.0 + got[0]) + got[1]) + got[2]
                assert tot == v or (np.isnan(tot) and np.isnan(v))


def test_free_and_reinitialise(tk):
    """tkb_free releases every buffer; the next constructor call starts from scratch and gives the same numbers"""
    xis, Q2 = np.array([0.2, 0.6]), np.array([4.0, 9.0])
    tk.matrices.initialize_kernels(10, xis, grid_type=2)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    before = tk.matrices.matrix_VSG().clone()
    L = tk._lib.lib()
    assert L.tkb_free() == 0
    assert L.tkb_get_nx() == 0 and L.tkb_get_nq2() == 0
    with pytest.raises(Exception):
        tk.matrices.initialize_evolution_matrices(Q2, nlo=True)            # kernels are gone: call order contract
    tk.matrices.initialize_kernels(10, xis, grid_type=2)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    assert torch.equal(tk.matrices.matrix_VSG(), before)


def test_cache_round_trip_feeds_the_application(tk, tmp_path):
    """matrices saved from the device, loaded back (whole and as a xi block) and applied"""
    from tiktaalik_b200 import cache
    nx, xis, Q2 = 12, np.array([0.1, 0.3, 0.5, 0.7]), tk.matrices.Q2space(4.0, 20.0, 3)
    tk.matrices.initialize_kernels(nx, xis, grid_type=1)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    mats = {"VNS_plus": tk.matrices.matrix_VNS(1), "VNS_minus": tk.matrices.matrix_VNS(-1),
            "VSG": tk.matrices.matrix_VSG(), "ASG": tk.matrices.matrix_ASG()}
    cache.save_matrices(str(tmp_path / "m"), mats, nx, xis, Q2, nlo=True, grid_type=1)
    meta, back = cache.load_matrices(str(tmp_path / "m"))
    assert meta["nx"] == nx and back["VSG"].is_cuda
    for k in mats:
        assert torch.equal(back[k], mats[k])
    g = torch.rand((2 * nx, 4, 3), dtype=torch.float64, device="cuda")
    assert torch.equal(tk.matrices.apply_matrix(back["VSG"], g), tk.matrices.apply_matrix(mats["VSG"], g))
    _, part = cache.load_matrices(str(tmp_path / "m"), xi_slice=slice(1, 3), types=["VSG"])
    out = tk.matrices.apply_matrix(part["VSG"], g[:, 1:3].contiguous())
    assert torch.equal(out, tk.matrices.apply_matrix(mats["VSG"], g)[:, 1:3])
