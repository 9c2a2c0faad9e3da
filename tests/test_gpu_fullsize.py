"""Full-size checks (nx = 160 / 320, grid 2, xi down to 2e-5, NLO: the shapes BASELINE.json names).  First through
size-independent properties on 8 xi (identity at the first Q2, LO plus == minus, A/V swap, composition of Q2 steps,
linearity of the application, momentum conservation of V-type LO evolution, a spot check of a few rows); then, at the
end of the file, FULL parity against the oracle's factored assembly at the benchmark's own shapes: every raw kernel
array for nfl = 3, 4, 5 bit for bit, every QAGS instance, the NLO evolution matrices within 1e-9 / 1e-12 max."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NX = 160
XI = np.geomspace(2e-5, 0.5, 8)


@pytest.fixture(scope="module")
def built(tk):
    tk.matrices.initialize_kernels(NX, XI, grid_type=2)
    return tk


def test_identity_swap_and_lo_degeneracy(built):
    tk = built
    Q2 = tk.matrices.Q2space(4.0, 10.0, 4)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=False)
    M = tk.matrices.matrix_VNS()
    eye = torch.eye(NX, dtype=torch.float64, device="cuda")
    assert all(torch.equal(M[:, :, k, 0], eye) for k in range(XI.shape[0]))
    assert torch.equal(tk.matrices.matrix_VNS(1), tk.matrices.matrix_VNS(-1))
    S = tk.matrices.matrix_VSG()
    assert torch.equal(S[:, :, 0, 0], torch.eye(2 * NX, dtype=torch.float64, device="cuda"))
    tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
    assert torch.equal(tk.matrices.matrix_ANS(1), tk.matrices.matrix_VNS(-1))
    assert torch.equal(tk.matrices.matrix_ANS(-1), tk.matrices.matrix_VNS(1))
    assert not torch.equal(tk.matrices.matrix_VNS(1), tk.matrices.matrix_VNS(-1))
    assert bool(torch.isfinite(tk.matrices.matrix_VSG()).all())


def test_q2_steps_compose(built):
    """M(Q0->Q2) = M(Q1->Q2) M(Q0->Q1) exactly as the stepping defines it (matevo.f90:521-523)"""
    tk = built
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 6.0, 10.0]), nlo=True)
    full = tk.matrices.matrix_VSG()[:, :, :, 2].clone()
    first = tk.matrices.matrix_VSG()[:, :, :, 1].clone()
    tk.matrices.initialize_evolution_matrices(np.array([6.0, 10.0]), nlo=True)
    second = tk.matrices.matrix_VSG()[:, :, :, 1]
    prod = torch.einsum('ijk,jlk->ilk', second, first)
    scale = full.abs().amax()
    assert float((prod - full).abs().amax() / scale) < 1e-12


def test_application_linearity_and_einsum(built):
    tk = built
    tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 10.0, 3), nlo=True)
    M = tk.matrices.matrix_VSG()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    for S in (1, 5, 64):
        g1 = torch.rand((2 * NX, XI.shape[0], S), generator=gen, dtype=torch.float64, device="cuda")
        g2 = torch.rand((2 * NX, XI.shape[0], S), generator=gen, dtype=torch.float64, device="cuda")
        o1, o2 = tk.matrices.apply_matrix(M, g1), tk.matrices.apply_matrix(M, g2)
        o12 = tk.matrices.apply_matrix(M, 2.0 * g1 - 3.0 * g2)
        ref = torch.einsum('ijkl,jks->iksl', M, g1)
        scale = ref.abs().amax()
        assert float((o1 - ref).abs().amax() / scale) < 1e-12
        assert float((o12 - (2.0 * o1 - 3.0 * o2)).abs().amax() / scale) < 1e-12


def test_lo_momentum_conservation(built, tko):
    """V-type LO evolution conserves int x Sigma + int G on a smooth input (discretisation accuracy)"""
    tk = built
    tk.matrices.initialize_evolution_matrices(np.array([4.0, 100.0]), nlo=False)
    M = tk.matrices.matrix_VSG().cpu().numpy()
    L = tko.lib()
    for k in (4, 7):                                           # xi = 6.5e-3 and 0.5
        xi = XI[k]
        x = tko.grid_nodes(NX, xi, 2)
        eta = np.array([(2 * i - 1) / NX - 1 for i in range(1, NX + 1)])
        w = np.array([L.tko_push_jacob(float(e), float(xi), NX, 2) for e in eta]) * (2.0 / NX)
        Hodd = (1 - x ** 2) ** 3 * x
        Hev = (1 - x ** 2) ** 3
        sg = np.concatenate([Hodd, Hev])
        mom0 = w @ (x * Hodd) + w @ Hev
        out = M[:, :, k, 1] @ sg
        mom1 = w @ (x * out[:NX]) + w @ out[NX:]
        assert abs(mom1 - mom0) < 2e-3 * abs(mom0), (k, mom0, mom1)
        assert np.abs(out - sg).max() > 1e-2                   # it did evolve


def test_rows_against_oracle_at_full_size(built, tko):
    """a few QAGS constants and kernel values at nx=160, xi = 2e-5 against the oracle (bit for bit)"""
    tk = built
    L = tk._lib.lib()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    res = tk._lib.view_tensor(a.value, (XI.shape[0], NX, 7, 3)).cpu().numpy()
    last = tk._lib.view_tensor(b.value, (XI.shape[0], NX, 7, 3), torch.int32).cpu().numpy()
    for k, i in ((0, 3), (0, 80), (0, 131), (5, 60), (7, 159)):
        x = tko.grid_nodes(NX, XI[k], 2)[i]
        for cst in (0, 3, 4, 6):
            v, l, _ = tko.cst_integral(cst, x, XI[k])
            got = res[k, i, cst]
            tot = ((0.0 + got[0]) + got[1]) + got[2]
            assert list(last[k, i, cst]) == l and (tot == v or (np.isnan(tot) and np.isnan(v))), (k, i, cst)
    assert int(last.max()) == 500                              # the V-type gluon constant hits QUADPACK's limit at xi = 2e-5


# --------------------------------------------------------------------------------------------
# Full-matrix parity at the benchmark's own shapes (matevo.f90:283-480 assembly, :485-673 stepping): every raw kernel
# array for nfl = 3, 4, 5 bit for bit, every QAGS instance's (result, last, ifail), and the NLO evolution matrices
# within the north-star tolerance -- at nx = 160 (C3/C4) and nx = 320 (C5), grid 2, including xi = 2e-5.

HEADLINE = [(160, [2e-5, 1e-3, 0.3]), (320, [2e-5, 0.45])]


def _full_parity(tk, tko, nx, xis, variant=(0, 0), with_qags=True, with_matrices=True):
    from test_gpu_parity import close, _raw_gpu, _qags_gpu
    xis = np.array(xis)
    f90 = tk.f90wrap.matrices.dummy
    try:
        f90.set_variant(*variant)
        tko.set_variant(*variant)
        tk.matrices.initialize_kernels(nx, xis, grid_type=2)
        tko.initialize_kernels(nx, xis, 2, mode=1)
        if with_qags:
            res, last, ifail = _qags_gpu(tk)
            for k, s in enumerate(xis):
                xs = tko.grid_nodes(nx, s, 2)
                for i in range(nx):
                    for c in range(7):
                        v, l, e = tko.cst_integral(c, xs[i], s)
                        assert list(last[k, i, c]) == l and list(ifail[k, i, c]) == e, (k, i, c, last[k, i, c], l)
                        got = res[k, i, c]
                        tot = ((0.0 + got[0]) + got[1]) + got[2]
                        assert tot == v or (np.isnan(tot) and np.isnan(v)), (k, i, c, tot, v)
        for nfl in (3, 4, 5):
            for which, name in enumerate(tko.RAW_NAMES):
                got = _raw_gpu(tk, which, nfl)
                ref = tko.raw_kernel(name, nfl)
                assert np.array_equal(got, ref, equal_nan=True), "%s nfl=%d nx=%d not bit-identical" % (name, nfl, nx)
        if with_matrices:
            Q2 = np.array([4.0, 10.0])
            tk.matrices.initialize_evolution_matrices(Q2, nlo=True)
            tko.initialize_evolution_matrices(Q2, nlo=True)
            for name, kw in (("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}), ("matrix_VSG", {}),
                             ("matrix_ASG", {})):
                close(getattr(tk.matrices, name)(**kw).cpu().numpy(), getattr(tko, name)(**kw))
    finally:
        f90.set_variant(0, 0)
        tko.set_variant(0, 0)


@pytest.mark.parametrize("nx,xis", HEADLINE)
def test_full_matrix_parity_at_headline_sizes(tk, tko, nx, xis):
    _full_parity(tk, tko, nx, xis)


@pytest.mark.parametrize("variant", [(1, 0), (0, 1)])
def test_ofast_variants_at_nx160(tk, tko, variant):
    """the two float32-division variants of an -Ofast reference build at the benchmark's grid"""
    _full_parity(tk, tko, 160, [2e-5, 0.3], variant=variant, with_qags=False, with_matrices=False)
