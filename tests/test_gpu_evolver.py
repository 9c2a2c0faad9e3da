"""GPU tests of the Evolver layer (reference: tiktaalik/evolution.py:19-548) against a numpy restatement of
the same pipeline driven by the oracle's matrices, on identical GK inputs (examples/examples.py:291-327)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=1e-9, afloor=1e-12):
    got = np.asarray(got)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = np.abs(ref).max()
    assert np.all(np.abs(got - ref) <= rtol * np.abs(ref) + afloor * scale), np.abs(got - ref).max() / scale


# ---- numpy restatement of evolution.py:435-540 on top of oracle matrices ----
def _tpls(nfl):
    M = np.zeros((nfl, nfl))
    M[0, :] = 1
    for i in range(1, nfl):
        M[i, 0:i] = 1
        M[i, i] = -i
    return M


def _pad(g, n):
    return np.concatenate((g, np.zeros((n,) + g.shape[1:])), axis=0)


def _evolve_fixed(g, M_SG, M_NSp, M_NSm, nfl):
    extra = g.shape[0] - 1 - nfl
    work = g.copy()
    q = work[1:nfl + 1]
    qmin = q + np.flip(q, axis=1)
    qpls = q - np.flip(q, axis=1)
    T = _tpls(nfl)
    st = np.einsum('ij,j...->i...', T, qpls)
    sg = np.concatenate((st[0], work[0]), axis=0)
    sg_f = np.einsum('ijkq,jk...->ik...q', M_SG, sg)
    t_f = np.einsum('ijkq,fjk...->fik...q', M_NSp, st[1:])
    qm_f = np.einsum('ijkq,fjk...->fik...q', M_NSm, qmin)
    nx = t_f.shape[1]
    st_f = np.concatenate((sg_f[None, 0:nx], t_f), axis=0)
    qp_f = np.einsum('ij,j...->i...', np.linalg.inv(T), st_f)
    q_f = (qp_f + qm_f) / 2
    out = np.concatenate((sg_f[None, nx:2 * nx], q_f), axis=0)
    return _pad(out, extra)


def _oracle_evolve_gpds(tko, nx, xi, grid_type, Q2, gpd, gpd_type, nlo, mc2, mb2, eps=1e-6):
    """Evolver.evolveGPDs (evolution.py:121-213) with oracle matrices"""
    tko.initialize_kernels(nx, xi, grid_type, mode=1)
    new = _pad(gpd[..., None], max(0, 5 - (gpd.shape[0] - 1)))
    pass_c = Q2[0] < mc2 - eps and Q2[-1] > mc2 + eps
    pass_b = Q2[0] < mb2 - eps and Q2[-1] > mb2 + eps
    segs = []
    if pass_c:
        segs.append((Q2[Q2 <= mc2], 3))
    if pass_b:
        segs.append((Q2[(mc2 <= Q2) & (Q2 <= mb2)], 4))
    last = Q2[mb2 <= Q2] if pass_b else (Q2[mc2 <= Q2] if pass_c else Q2)
    nfl_last = 3 if Q2[-1] < mc2 else (4 if Q2[-1] < mb2 else 5)
    segs.append((last, nfl_last))
    for sub, nfl in segs:
        tko.initialize_evolution_matrices(sub, nlo=nlo)
        if gpd_type == 'V':
            M_SG, M_p, M_m = tko.matrix_VSG(), tko.matrix_VNS(), (tko.matrix_ANS() if nlo else tko.matrix_VNS())
        else:
            M_SG, M_p, M_m = tko.matrix_ASG(), tko.matrix_ANS(), (tko.matrix_VNS() if nlo else tko.matrix_ANS())
        added = _evolve_fixed(new[..., -1], M_SG, M_p, M_m, nfl)
        new = np.concatenate((new, added[..., 1:]), axis=-1)
    return new


@pytest.mark.parametrize("gpd_type,nlo", [("V", False), ("V", True), ("A", True)])
def test_evolver_evolve_gpds_through_thresholds(tk, tko, gpd_type, nlo):
    nx, xi = 16, np.array([0.2, 0.5])
    Q2i, Q2f, nQ2 = 1.2, 30.0, 7                       # crosses mc2 and mb2: three segments
    ev = tk.Evolver(nx=nx, xi=xi, nQ2=nQ2, Q2i=Q2i, Q2f=Q2f, grid_type=1)
    assert ev.pass_charm and ev.pass_bottom and ev.Q2.shape == (nQ2,)
    x = tko.pixelspace(nx)
    t = np.array([0.0, -0.3])
    gpd = np.stack([tko.gk(n, x, xi, t) for n in ("Hg", "Hu", "Hd", "Hs")], axis=0)        # (4, nx, nxi, nt)
    got = ev.evolveGPDs(gpd, gpd_type=gpd_type, nlo=nlo)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert tuple(got.shape) == (6, nx, 2, 2, nQ2)
    ref = _oracle_evolve_gpds(tko, nx, xi, 1, ev.Q2, gpd, gpd_type, nlo, tk.pars.mc2, tk.pars.mb2)
    _close(got.cpu().numpy(), ref)
    # the first Q2 slice is the (padded) input; heavy flavours start at zero
    assert np.array_equal(got[:4, ..., 0].cpu().numpy(), gpd) and float(got[4:, ..., 0].abs().max()) == 0.0


def test_evolver_accepts_a_seven_row_grid(tk, tko):
    """gluon + six quark rows (g, u, d, s, c, b, t): the reference keeps the extra row and zero-pads it after the
    first Q2 point through add_pad (evolution.py:150-153, 495-514); ADVICE round 1: the library refused Fout > 6"""
    nx, xi = 16, np.array([0.3, 0.6])
    ev = tk.Evolver(nx=nx, xi=xi, nQ2=5, Q2i=1.2, Q2f=30.0, grid_type=1)
    x = tko.pixelspace(nx)
    base = np.stack([tko.gk(n, x, xi, [0.0]) for n in ("Hg", "Hu", "Hd", "Hs")], axis=0)[..., 0]       # (4, nx, nxi)
    rng = np.random.default_rng(5)
    gpd = np.concatenate((base, 0.1 * rng.standard_normal((3, nx, 2))), axis=0)                       # 7 rows
    got = ev.evolveGPDs(gpd, gpd_type='V', nlo=True)
    assert tuple(got.shape) == (7, nx, 2, 5)
    ref = _oracle_evolve_gpds(tko, nx, xi, 1, ev.Q2, gpd, 'V', True, tk.pars.mc2, tk.pars.mb2)
    _close(got.cpu().numpy(), ref)


def test_evolver_single_sector_methods(tk, tko):
    nx, xi = 16, np.array([0.3])
    ev = tk.Evolver(nx=nx, xi=xi, nQ2=5, Q2i=4.0, Q2f=10.0)
    assert not ev.pass_charm and not ev.pass_bottom
    x = tko.pixelspace(nx)
    umin = (tko.gk("Hu", x, xi, [0.0]) + tko.gk("Hu", -x, xi, [0.0]))[:, :, 0]       # (nx, nxi)
    sing = tko.H_singlet(x, xi, [0.0])[:, :, 0]
    glue = tko.gk("Hg", x, xi, [0.0])[:, :, 0]
    tko.initialize_kernels(nx, xi, 1, mode=1)
    for nlo in (False, True):
        tko.initialize_evolution_matrices(ev.Q2, nlo=nlo)
        for ns_type, M in ((1, tko.matrix_VNS(1)), (-1, tko.matrix_VNS(-1))):
            got = ev.evolveNS(umin, gpd_type='V', nlo=nlo, ns_type=ns_type)
            assert tuple(got.shape) == (nx, 1, 5)
            _close(got.cpu().numpy(), np.einsum('ijkl,jk...->ik...l', M, umin))
        # A-type minus is V-type plus (evolution.py:318-320)
        got = ev.evolveNS(umin, gpd_type='A', nlo=nlo, ns_type=-1)
        _close(got.cpu().numpy(), np.einsum('ijkl,jk...->ik...l', tko.matrix_VNS(1), umin))
        sg = np.concatenate((sing, glue), axis=0)
        for typ, M in (("V", tko.matrix_VSG()), ("A", tko.matrix_ASG())):
            if typ == "V" and nlo:
                continue                                  # reference selector bug, tested below
            s, g = ev.evolveSinglet(sing, glue, gpd_type=typ, nlo=nlo)
            ref = np.einsum('ijkl,jk...->ik...l', M, sg)
            _close(s.cpu().numpy(), ref[:nx])
            _close(g.cpu().numpy(), ref[nx:])
    # evolution.py:273-280: V / NLO never selects the final-segment matrix; with no threshold passed the
    # reference dies with UnboundLocalError
    with pytest.raises(UnboundLocalError):
        ev.evolveSinglet(sing, glue, gpd_type='V', nlo=True)
    import tiktaalik_b200.evolution as evo
    evo.REFERENCE_BUGS = False
    try:
        tko.initialize_evolution_matrices(ev.Q2, nlo=True)
        s, g = ev.evolveSinglet(sing, glue, gpd_type='V', nlo=True)
        ref = np.einsum('ijkl,jk...->ik...l', tko.matrix_VSG(), np.concatenate((sing, glue), axis=0))
        _close(s.cpu().numpy(), ref[:nx])
    finally:
        evo.REFERENCE_BUGS = True


def test_evolver_api_contract(tk):
    with pytest.raises(ValueError):
        tk.Evolver(nx=15)
    with pytest.raises(AssertionError):
        tk.matrices.initialize_kernels(7, 0.5)
    with pytest.raises(AssertionError):
        tk.matrices.initialize_evolution_matrices(np.array([4.0]))
    tk.matrices.initialize_kernels(12, 0.5)
    with pytest.raises(RuntimeError, match="not initialised"):
        tk.matrices.matrix_VNS()                         # make_kernels frees the matrices (matevo.f90:70)
    tk.matrices.initialize_evolution_matrices(np.array([4, 8, 12, 16]))          # integer Q2 as in examples.py:62
    M = tk.matrices.matrix_VNS()
    assert tuple(M.shape) == (12, 12, 1, 4)
    M[0, 0, 0, 0] = 7.0                                   # getters return private copies (f2py intent(out))
    assert float(tk.matrices.matrix_VNS()[0, 0, 0, 0]) == 1.0
    assert tuple(tk.matrices.matrix_VSG().shape) == (24, 24, 1, 4)
    assert tuple(tk.matrices.kernel_VGG().shape) == (12, 12, 1)
    with pytest.raises(AssertionError):
        tk.matrices.kernel_VQQ(nfl=6)


def test_gk_model_on_device_bit_identical(tk, tko):
    """tiktaalik.model API on the device (model/gk.f90:75-241) against the oracle's restatement"""
    x = tko.pixelspace(40)
    xi = np.array([0.05, 0.3, 0.7])
    t = np.array([0.0, -0.25, -1.0])
    for name in ("Hu", "Hd", "Hs", "Hg", "Hu_val", "Hd_val", "Hu_sea", "Hd_sea"):
        got = getattr(tk.model, name)(x, xi, t)
        assert got.is_cuda and tuple(got.shape) == (40, 3, 3)
        ref = tko.gk(name, x, xi, t)
        assert np.array_equal(got.cpu().numpy(), ref, equal_nan=True), name
    s = tk.model.H_singlet(x, xi, t).cpu().numpy()
    assert np.array_equal(s, tko.H_singlet(x, xi, t))
    # scalars are accepted like in the reference (examples/examples.py:46)
    assert tuple(tk.model.Hu(x, 0.5, t=0).shape) == (40, 1, 1)


@pytest.mark.parametrize("nfl,F,extra", [(3, 4, (5,)), (4, 6, (2, 3)), (5, 4, ()), (4, 3, (40,))])
def test_fused_fixed_flavor_step_equals_the_unfused_pipeline(tk, nfl, F, extra):
    """csrc/evolve.cu against the reference-shaped sequence of tensor operations (evolution.py:435-514)"""
    nx, xis = 12, np.array([0.1, 0.3, 0.6])
    tk.matrices.initialize_kernels(nx, xis, grid_type=2)
    tk.matrices.initialize_evolution_matrices(tk.matrices.Q2space(4.0, 40.0, 5), nlo=True)
    M_SG, M_p, M_m = tk.matrices.matrix_VSG(), tk.matrices.matrix_VNS(1), tk.matrices.matrix_VNS(-1)
    gen = torch.Generator(device="cuda").manual_seed(11)
    g = torch.rand((F, nx, len(xis)) + extra, generator=gen, dtype=torch.float64, device="cuda") - 0.3
    a = tk.evolution.evolve_fixed_flavor(g, M_SG, M_p, M_m, nfl=nfl)
    b = tk.evolution.evolve_fixed_flavor_unfused(g, M_SG, M_p, M_m, nfl=nfl)
    assert tuple(a.shape) == tuple(b.shape)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-13 * float(b.abs().max()))
