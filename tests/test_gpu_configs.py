"""The configurations BASELINE.json names, at their real grid sizes, against the oracle where it finishes in
seconds (C1, C2: nx = 80) and through exact properties where it does not (C4: nx = 160, 128 xi, 16 Q2 applied to
sampled GPDs; C5: nx = 320).  C3 (nx = 160, xi down to 2e-5) is tests/test_gpu_fullsize.py."""
import numpy as np
import pytest
import torch

from test_gpu_parity import close, _raw_gpu

pytestmark = pytest.mark.gpu

NAMES = (("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}), ("matrix_ASG", {}))


def test_config_c1_single_xi_nx80_with_gk_input(tk, tko):
    """examples/examples.py:37-99: nx = 80 linear grid, xi = 0.5, Q2 4 -> 10 GeV^2, GK model input at t = 0"""
    nx, xi, Q2 = 80, 0.5, np.array([4.0, 10.0])
    tk.matrices.initialize_kernels(nx, xi, grid_type=1)
    tko.initialize_kernels(nx, [xi], 1, mode=1)
    for which, name in enumerate(tko.RAW_NAMES):
        assert np.array_equal(_raw_gpu(tk, which, 4), tko.raw_kernel(name, 4), equal_nan=True), name
    x = tko.pixelspace(nx)
    xd = tk.matrices.pixelspace(nx)
    assert np.array_equal(np.asarray(xd.cpu() if isinstance(xd, torch.Tensor) else xd), x)
    umin0 = tko.gk("Hu", x, [xi], [0.0]) + tko.gk("Hu", -x, [xi], [0.0])
    sg0 = np.concatenate((tko.H_singlet(x, [xi], [0.0]), tko.gk("Hg", x, [xi], [0.0])), axis=0)
    for nlo in (False, True):
        tk.matrices.initialize_evolution_matrices(Q2, nlo=nlo)
        tko.initialize_evolution_matrices(Q2, nlo=nlo)
        for name, kw in NAMES:
            close(getattr(tk.matrices, name)(**kw).cpu().numpy(), getattr(tko, name)(**kw))
        close(tk.matrices.apply_matrix(tk.matrices.matrix_VNS(), umin0).cpu().numpy(),
              np.einsum('ijkl,jk...->ik...l', tko.matrix_VNS(), umin0))
        close(tk.matrices.apply_matrix(tk.matrices.matrix_VSG(), sg0).cpu().numpy(),
              np.einsum('ijkl,jk...->ik...l', tko.matrix_VSG(), sg0))


@pytest.mark.parametrize("q2max", [10.0, 100.0])
def test_config_c2_64_xi_20_q2_lo(tk, tko, q2max):
    """nx = 80, 64 xi in [0.05, 0.9], 20 Q2 points (and the variant that crosses mb2: nfl 4 -> 5), LO, GK u, d, s, g
    input; the GPU does all 64 xi, the oracle every fourth of them"""
    nx = 80
    xis = np.linspace(0.05, 0.9, 64)
    sub = np.arange(0, 64, 4)
    Q2 = tk.matrices.Q2space(4.0, q2max, 20)
    tk.matrices.initialize_kernels(nx, xis, grid_type=1)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=False)
    tko.initialize_kernels(nx, xis[sub], 1, mode=1)
    tko.initialize_evolution_matrices(Q2, nlo=False)
    x = tko.pixelspace(nx)
    for name, kw in NAMES:
        got = getattr(tk.matrices, name)(**kw)
        assert tuple(got.shape)[2:] == (64, 20)
        close(got[:, :, torch.as_tensor(sub, device=got.device), :].cpu().numpy(), getattr(tko, name)(**kw))
    # singlet + gluon of the GK model at t = 0 through the application kernel, all xi at once
    sg0 = np.concatenate((tko.H_singlet(x, xis, [0.0]), tko.gk("Hg", x, xis, [0.0])), axis=0)[:, :, 0]
    got = tk.matrices.apply_matrix(tk.matrices.matrix_VSG(), sg0).cpu().numpy()
    ref = np.einsum('ijkl,jk->ikl', tko.matrix_VSG(), sg0[:, sub])
    close(got[:, sub, :], ref)


def test_config_c4_application_at_full_matrix_size(tk):
    """nx = 160, 128 xi in [1e-3, 0.9] (grid 2), 16 Q2, NLO V-type matrices applied to samples of the 4-parameter
    family g(x) = N u^-a (1-u)^b (1 + c x), u = |x| (seed 1234): every regime of the application kernel against
    einsum on the same matrices"""
    nx, nxi, nq = 160, 128, 16
    xis = np.geomspace(1e-3, 0.9, nxi)
    tk.matrices.initialize_kernels(nx, xis, grid_type=2)
    tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, nq), nlo=True)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    S = 96
    a = torch.rand(S, generator=gen, device="cuda", dtype=torch.float64) * 0.5
    b = 2 + 3 * torch.rand(S, generator=gen, device="cuda", dtype=torch.float64)
    c = 2 * torch.rand(S, generator=gen, device="cuda", dtype=torch.float64) - 1
    N = 0.5 + 1.5 * torch.rand(S, generator=gen, device="cuda", dtype=torch.float64)
    xg = torch.as_tensor(np.asarray(tk.matrices.pixelspace(nx)), dtype=torch.float64, device="cuda")
    u = xg.abs()[:, None]
    g1 = N * u ** (-a) * (1 - u) ** b * (1 + c * xg[:, None])                       # (nx, S)
    g_ns = g1[:, None, :].expand(nx, nxi, S).contiguous()
    M = tk.matrices.matrix_VNS()
    for s in (1, 4, 7, 24, 96):
        out = tk.matrices.apply_matrix(M, g_ns[:, :, :s].contiguous())
        ref = torch.einsum("ijkl,jks->iksl", M, g_ns[:, :, :s])
        assert tuple(out.shape) == (nx, nxi, s, nq)
        assert torch.allclose(out, ref, rtol=1e-11, atol=1e-12 * float(ref.abs().max())), s
    Msg = tk.matrices.matrix_VSG()
    g_sg = torch.cat((g_ns, 0.3 * g_ns.flip(0)), dim=0)
    for s in (2, 16, 80):
        out = tk.matrices.apply_matrix(Msg, g_sg[:, :, :s].contiguous())
        ref = torch.einsum("ijkl,jks->iksl", Msg, g_sg[:, :, :s])
        assert torch.allclose(out, ref, rtol=1e-11, atol=1e-12 * float(ref.abs().max())), s
    # linearity in the samples and the identity at the first Q2 point
    out = tk.matrices.apply_matrix(M, g_ns[:, :, :8].contiguous())
    assert torch.equal(out[..., 0], g_ns[:, :, :8])
    comb = tk.matrices.apply_matrix(M, (g_ns[:, :, :4] + 2.0 * g_ns[:, :, 4:8]).contiguous())
    assert torch.allclose(comb, out[:, :, :4] + 2.0 * out[:, :, 4:8], rtol=1e-12, atol=1e-12 * float(out.abs().max()))


def test_sample_blocks_are_read_in_place(tk):
    """a block g_all[:, :, s0:s1] of a larger sample array goes to the library with its strides (tkb_apply_strided)
    and gives the bits of the contiguous copy, for the warp-per-row kernel (m < 40) and the GEMM"""
    for nx in (12, 40):
        xis = np.array([0.1, 0.4, 0.8])
        tk.matrices.initialize_kernels(nx, xis, grid_type=1)
        tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 30.0, 3), nlo=True)
        M = tk.matrices.matrix_VSG()
        g_all = torch.rand((2 * nx, 3, 37), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        for s0, s1 in ((0, 37), (3, 4), (5, 12), (20, 37)):
            blk = g_all[:, :, s0:s1]
            assert not blk.is_contiguous() or s1 - s0 == 37
            assert torch.equal(tk.matrices.apply_matrix(M, blk), tk.matrices.apply_matrix(M, blk.contiguous())), (nx, s0, s1)
        parts = []
        for s0, s1, o in tk.matrices.apply_matrix_chunks(M, g_all, chunk=8):
            assert torch.equal(o, tk.matrices.apply_matrix(M, g_all[:, :, s0:s1].contiguous())), (nx, s0)
            parts.append(o.clone())
        whole = tk.matrices.apply_matrix(M, g_all)      # another block width may sum in another order
        assert torch.allclose(torch.cat(parts, dim=2), whole, rtol=1e-12, atol=1e-13 * float(whole.abs().max()))
        with pytest.raises(ValueError):
            tk.matrices.f90src.apply_wrap(M, g_all, out=torch.empty(3, device="cuda", dtype=torch.float64))


def test_config_c4_all_100000_samples_in_blocks(tk):
    """BASELINE config 4 at its full sample count: nx = 160, 128 xi, 16 Q2, 10^5 samples (16 GB of input; the
    262 GB result never exists as a whole) applied 2048 at a time.  Exact per block: the first Q2 point returns
    the input.  Over the whole run: the sum over samples of the results equals the result of the summed samples."""
    nx, nxi, nq, S = 160, 128, 16, 100000
    xis = np.geomspace(1e-3, 0.9, nxi)
    tk.matrices.initialize_kernels(nx, xis, grid_type=2)
    tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, nq), nlo=True)
    M = tk.matrices.matrix_VNS()
    gen = torch.Generator(device="cuda").manual_seed(99)
    g_all = torch.rand((nx, nxi, S), generator=gen, dtype=torch.float64, device="cuda")
    acc = torch.zeros((nx, nxi, nq), dtype=torch.float64, device="cuda")
    seen = 0
    for s0, s1, out in tk.matrices.apply_matrix_chunks(M, g_all, chunk=2048):
        assert tuple(out.shape) == (nx, nxi, s1 - s0, nq)
        assert torch.equal(out[..., 0], g_all[:, :, s0:s1])
        acc += out.sum(dim=2)
        seen += s1 - s0
    assert seen == S
    ref = tk.matrices.apply_matrix(M, g_all.sum(dim=2, keepdim=True))[:, :, 0, :]
    assert torch.allclose(acc, ref, rtol=1e-9, atol=1e-12 * float(ref.abs().max()))
    # one block in the middle against einsum on the same matrices
    s0 = 51200
    out = tk.matrices.apply_matrix(M, g_all[:, :, s0:s0 + 64])
    ref = torch.einsum("ijkl,jks->iksl", M, g_all[:, :, s0:s0 + 64])
    assert torch.allclose(out, ref, rtol=1e-11, atol=1e-12 * float(ref.abs().max()))


def test_config_c5_shard_shape_nx320(tk, tko):
    """nx = 320, grid 2 (one GPU's block of the 4096-xi sweep, here 3 xi): exact identities, a few QAGS constants
    and kernel rows against the oracle"""
    import ctypes as C
    nx = 320
    xis = np.array([2e-5, 3e-3, 0.45])
    tk.shard.f90src.make_kernels_shard(nx, xis, 2, 0, 0)
    L = tk._lib.lib()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    res = tk._lib.view_tensor(a.value, (3, nx, 7, 3)).cpu().numpy()
    last = tk._lib.view_tensor(b.value, (3, nx, 7, 3), torch.int32).cpu().numpy()
    for k, i in ((0, 7), (0, 200), (1, 160), (2, 319)):
        xnode = tko.grid_nodes(nx, xis[k], 2)[i]
        for cst in (1, 3, 5):
            v, l, _ = tko.cst_integral(cst, xnode, xis[k])
            got = res[k, i, cst]
            tot = ((0.0 + got[0]) + got[1]) + got[2]
            assert list(last[k, i, cst]) == l and (tot == v or (np.isnan(tot) and np.isnan(v))), (k, i, cst)
    Q2 = np.array([4.0, 10.0])
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    M = tk.matrices.matrix_VNS()
    assert tuple(M.shape) == (nx, nx, 3, 2)
    eye = torch.eye(nx, dtype=torch.float64, device="cuda")
    assert all(torch.equal(M[:, :, k, 0], eye) for k in range(3))
    assert torch.equal(tk.matrices.matrix_ANS(1), tk.matrices.matrix_VNS(-1))
    S = tk.matrices.matrix_VSG()
    assert tuple(S.shape) == (2 * nx, 2 * nx, 3, 2) and bool(torch.isfinite(S).all())
    # LO kernel rows of the first xi against the oracle's reference-shaped assembly of the same row
    tko.initialize_kernels(nx, xis[2:3], 2, mode=1)
    got = _raw_gpu(tk, 0, 4)[:, :, 2]
    ref = tko.raw_kernel("K_NS_0", 4)[:, :, 0]
    assert np.array_equal(got, ref, equal_nan=True)
