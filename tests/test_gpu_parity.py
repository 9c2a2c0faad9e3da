"""GPU parity tests: the CUDA path (through the C ABI / the tiktaalik-style Python API) against the
CPU oracle on the same seeded inputs.  Bars (north_star / SURVEY.md 8d):
  * transcendental primitives and kernel functions: bit-identical to the oracle
    (this is what makes 1e-9 reachable at small xi);
  * kernel matrices, QAGS constants: relative <= 1e-9 with absolute floor 1e-12 * max|entry|,
    and the same QAGS path (last, ifail) per instance;
  * evolution matrices and evolved GPDs: relative <= 1e-9 (same floor).
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-9
AFLOOR = 1e-12


def close(got, ref, rtol=RTOL, afloor=AFLOOR):
    """|got-ref| <= rtol*|ref| + afloor*max|ref|, NaN patterns equal; returns worst normalised error"""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN pattern differs"
    fin = np.isfinite(ref)
    assert np.array_equal(got[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)])
    scale = np.abs(ref[fin]).max() if fin.any() else 1.0
    err = np.abs(got[fin] - ref[fin])
    tol = rtol * np.abs(ref[fin]) + afloor * scale
    bad = err > tol
    assert not bad.any(), "max err %.3e (scale %.3e), %d entries out of tolerance" % (err.max(), scale, bad.sum())
    return float((err / (np.abs(ref[fin]) + afloor * scale / rtol)).max()) if err.size else 0.0


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


# --------------------------------------------------------------------------------------------
# primitives

def _eval_scalar(tk, which, a, b=None):
    L = tk._lib.lib()
    a_d = dev(a)
    b_d = dev(b if b is not None else a)
    out = torch.empty_like(a_d)
    tk._lib.check(L.tkb_eval_scalar(which, a_d.numel(), tk._lib.ptr(a_d), tk._lib.ptr(b_d), tk._lib.ptr(out),
                                    tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_log_pow_bit_identical_to_oracle(tk, tko):
    rng = np.random.default_rng(1)
    L = tko.lib()
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), 1 + rng.uniform(-1e-3, 1e-3, 5000),
                        rng.uniform(0, 2, 20000), [5e-324, 1e-310, 1.0, 2.0, 0.5, 1e308, 0.0, -1.0, np.inf]])
    got = _eval_scalar(tk, 0, x)
    ref = np.array([L.tko_log_x(float(v)) for v in x])
    assert np.array_equal(bits(got)[~np.isnan(ref)], bits(ref)[~np.isnan(ref)])
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    base = np.concatenate([rng.uniform(1e-12, 1e-2, 20000), rng.uniform(1e-10, 1.0, 5000),
                           np.exp(rng.uniform(-50, 50, 5000))])
    expo = np.concatenate([rng.uniform(0.0, 2.0, 25000), np.full(5000, 1.5)])
    got = _eval_scalar(tk, 1, base, expo)
    ref = np.array([L.tko_pow_x(float(b), float(e)) for b, e in zip(base, expo)])
    assert np.array_equal(bits(got), bits(ref))
    # the grid-type-2 node maps themselves, extreme exponents, and the grid maps' own logarithm
    nodes = (np.float32(2 * np.arange(1, 321) - 1) / np.float32(320) - np.float32(1)).astype(np.float64)
    xi = np.geomspace(2e-5, 0.9, 200)
    base = np.repeat(xi * xi, nodes.size)
    expo = np.tile(1.0 - np.abs(nodes), xi.size)
    base = np.concatenate([base, [1e-300, 1e300, 2.0, 2.0, 0.5, 1 + 2.0 ** -52, 5e-324, 10.0, 10.0]])
    expo = np.concatenate([expo, [1.05, 1.02, 1023.5, -1074.0, 1074.5, 2.0 ** 40, 0.5, 400.0, -400.0]])
    got = _eval_scalar(tk, 1, base, expo)
    ref = np.array([L.tko_pow_x(float(b), float(e)) for b, e in zip(base, expo)])
    assert np.array_equal(bits(got), bits(ref))
    got = _eval_scalar(tk, 3, x)
    ref = np.array([L.tko_log_cr_x(float(v)) for v in x])
    assert np.array_equal(bits(got)[~np.isnan(ref)], bits(ref)[~np.isnan(ref)])
    assert np.array_equal(np.isnan(got), np.isnan(ref))


def test_dilog_bit_identical_to_oracle(tk, tko):
    rng = np.random.default_rng(2)
    z = np.concatenate([rng.uniform(-30, 30, 40000), rng.uniform(-1.5, 2.5, 40000),
                        -np.exp(rng.uniform(0, 25, 5000)), np.exp(rng.uniform(0, 25, 5000)),
                        [-1.0, 0.0, 0.5, 1.0, 2.0]])
    edge = []
    for c in (-1.0, 0.0, 0.5, 1.0, 2.0):
        v = c
        for _ in range(40):
            v = np.nextafter(v, np.inf)
            edge.append(v)
        v = c
        for _ in range(40):
            v = np.nextafter(v, -np.inf)
            edge.append(v)
    z = np.concatenate([z, edge])
    got = _eval_scalar(tk, 2, z)
    ref = np.array([tko.dilog(v) for v in z])
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.parametrize("xi", [0.5, 0.05, 1e-3, 2e-5])
def test_all_kernel_functions_bit_identical(tk, tko, xi):
    """the 22 two-variable kernels (kernels_lo.f90, kernels_nlo.f90) at random points of all regions"""
    rng = np.random.default_rng(3)
    n = 1500
    x = rng.uniform(-1, 1, n)
    y = rng.uniform(-1, 1, n)
    # populate the ERBL region too when xi is small
    x[: n // 3] = rng.uniform(-1, 1, n // 3) * xi * 1.5
    y[n // 6: n // 2] = rng.uniform(-1, 1, n // 2 - n // 6) * xi * 1.5
    L = tk._lib.lib()
    out = torch.empty((n, 22), dtype=torch.float64, device="cuda")
    xd, yd, xid = dev(x), dev(y), dev(np.full(n, xi))
    tk._lib.check(L.tkb_eval_kernels(n, tk._lib.ptr(xd), tk._lib.ptr(yd), tk._lib.ptr(xid), tk._lib.ptr(out),
                                     tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = np.array([[tko.kernel(f, x[p], y[p], xi) for f in range(22)] for p in range(n)])
    same = bits(got) == bits(ref)
    both_nan = np.isnan(got) & np.isnan(ref)
    ok = same | both_nan | ((got == 0) & (ref == 0))
    assert ok.all(), "kernels not bit-identical: %s" % sorted(set(np.nonzero(~ok)[1].tolist()))


@pytest.mark.parametrize("xi", [0.3, 2e-3])
def test_qags_integrand_hook_bit_identical(tk, tko, xi):
    """tkb_eval_integrand: K(x,y,xi) - K(y,x,xi) of the seven plus-prescription constants (kernels_nlo.f90:71-77),
    the straight evaluation whose FP64 instruction count calibrates the K2 roofline"""
    rng = np.random.default_rng(9)
    n = 600
    x = rng.uniform(-1, 1, n)
    y = rng.uniform(-1, 1, n)
    x[: n // 3] *= xi * 1.5
    y[n // 6: n // 2] *= xi * 1.5
    L = tk._lib.lib()
    xd, yd, xid = dev(x), dev(y), dev(np.full(n, xi))
    for c, f in enumerate((7, 8, 9, 18, 19, 20, 21)):
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        tk._lib.check(L.tkb_eval_integrand(c, n, tk._lib.ptr(xd), tk._lib.ptr(yd), tk._lib.ptr(xid), tk._lib.ptr(out),
                                           tk._lib.stream_ptr()))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        ref = np.array([tko.kernel(f, x[p], y[p], xi) - tko.kernel(f, y[p], x[p], xi) for p in range(n)])
        ok = (bits(got) == bits(ref)) | (np.isnan(got) & np.isnan(ref)) | ((got == 0) & (ref == 0))
        assert ok.all(), (c, int((~ok).sum()))


def test_kernel_functions_at_singular_points(tk, tko):
    """x = y, y = +-xi, x = +-xi, zeros and end points: the reference evaluates pieces outside their support and
    multiplies by a vanishing step, so 0 * Inf = NaN, divisions by zero and log(0) must come out the same
    (kernels_nlo.f90:137-141): same NaN pattern, same infinities, same finite bits"""
    xi = 0.25
    vals = [-1.0, -0.6, -xi, -0.1, 0.0, 0.1, xi, 0.6, 1.0, 2 * xi - 1, 1 - 2 * xi]
    pts = [(a, b) for a in vals for b in vals]
    x = np.array([p[0] for p in pts])
    y = np.array([p[1] for p in pts])
    n = len(pts)
    L = tk._lib.lib()
    out = torch.empty((n, 22), dtype=torch.float64, device="cuda")
    xd, yd, xid = dev(x), dev(y), dev(np.full(n, xi))
    tk._lib.check(L.tkb_eval_kernels(n, tk._lib.ptr(xd), tk._lib.ptr(yd), tk._lib.ptr(xid), tk._lib.ptr(out),
                                     tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = np.array([[tko.kernel(f, x[p], y[p], xi) for f in range(22)] for p in range(n)])
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN pattern differs at %s" % np.argwhere(np.isnan(got) != np.isnan(ref))[:5]
    fin = ~np.isnan(ref)
    assert np.array_equal(got[fin], ref[fin]), np.argwhere((got != ref) & fin)[:5]
    assert np.isnan(ref).any() and np.isinf(ref[fin]).any()          # the points do exercise the singular cases


@pytest.mark.parametrize("c", range(8))
def test_skipped_tasks_are_exact_zeros_and_erbl_only_tasks_lose_nothing(tk, c):
    """(c = 7: all 22 kernels as the fixed-quadrature kernel evaluates them; `full` then counts non-zero values.)
    The assembly kernels skip half kernels of class 0 and evaluate only the ERBL pieces of class 1.  On random,
    near-singular (a few ulps from x = y, y = +-xi, 0, +-1) and exactly singular nodes: class 0 => the full evaluation
    is an exact zero, class 1 => the erbl-only evaluation equals the full one bit for bit (NaN for NaN)."""
    rng = np.random.default_rng(100 + c)
    pts = []
    for xi in (0.5, 0.05, 1e-3, 2e-5):
        n = 1500
        x = rng.uniform(-1, 1, n)
        y = rng.uniform(-1, 1, n)
        x[: n // 3] = rng.uniform(-1.5, 1.5, n // 3) * xi
        y[n // 6: n // 2] = rng.uniform(-1.5, 1.5, n // 2 - n // 6) * xi
        special = np.array([0.0, xi, -xi, 1.0, -1.0, 0.3, -0.3, 2 * xi, 1 - 2 * xi if xi < 0.5 else 0.1])
        near = np.concatenate([np.nextafter(special, 2.0), np.nextafter(special, -2.0), special,
                               special * (1 + 4e-16), special * (1 - 4e-16)])
        xs = np.concatenate([x, rng.choice(near, 600), rng.choice(near, 600), rng.uniform(-1, 1, 300)])
        ys = np.concatenate([y, rng.choice(near, 600), rng.uniform(-1, 1, 600), rng.choice(near, 300)])
        # y a few ulps from x as well
        xs = np.concatenate([xs, x[:300]])
        ys = np.concatenate([ys, np.nextafter(x[:300], rng.choice([-2.0, 2.0], 300))])
        pts.append((xs, ys, np.full(xs.shape[0], xi)))
    x = np.concatenate([p[0] for p in pts]); y = np.concatenate([p[1] for p in pts]); xi = np.concatenate([p[2] for p in pts])
    n = x.shape[0]
    L = tk._lib.lib()
    cls = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    full = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    erbl = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    xd, yd, xid = dev(x), dev(y), dev(xi)
    tk._lib.check(L.tkb_task_classes(c, n, tk._lib.ptr(xd), tk._lib.ptr(yd), tk._lib.ptr(xid), tk._lib.ptr(cls),
                                     tk._lib.ptr(full), tk._lib.ptr(erbl), tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    cls, full, erbl = cls.cpu().numpy(), full.cpu().numpy(), erbl.cpu().numpy()
    dead = cls == 0
    assert np.all(full[dead] == 0.0), "a skipped task is not an exact zero: %s" % np.argwhere(dead & (full != 0.0))[:5]
    if c == 7:
        assert dead.sum() > 1000 and (full[~dead] > 0).any()
        return
    one = cls == 1
    same = (bits(full) == bits(erbl)) | (np.isnan(full) & np.isnan(erbl)) | ((full == 0) & (erbl == 0))
    assert np.all(same[one]), "erbl-only evaluation differs: %s" % np.argwhere(one & ~same)[:5]
    assert dead.sum() > 1000 and one.sum() > 100 and (cls == 2).sum() > 1000       # all classes are exercised
    assert np.isnan(full).any()                                                     # and so are the singular nodes


# --------------------------------------------------------------------------------------------
# assembly

def _raw_gpu(tk, which, nfl):
    L = tk._lib.lib()
    n, nxi = L.tkb_get_nx(), L.tkb_get_nxi()
    m = n * (2 if which >= 3 else 1)
    addr = C.c_void_p()
    tk._lib.check(L.tkb_raw_kernel(which, nfl, C.byref(addr), tk._lib.stream_ptr()))
    torch.cuda.synchronize()
    t = tk._lib.view_tensor(addr.value, (nxi, m, m))
    return t.permute(1, 2, 0).cpu().numpy()              # (i, j, xi) like the oracle's Fortran order


def _qags_gpu(tk):
    L = tk._lib.lib()
    n, nxi = L.tkb_get_nx(), L.tkb_get_nxi()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    tk._lib.check(L.tkb_qags_info(C.byref(a), C.byref(b), C.byref(c)))
    res = tk._lib.view_tensor(a.value, (nxi, n, 7, 3)).cpu().numpy()
    last = tk._lib.view_tensor(b.value, (nxi, n, 7, 3), torch.int32).cpu().numpy()
    ifail = tk._lib.view_tensor(c.value, (nxi, n, 7, 3), torch.int32).cpu().numpy()
    return res, last, ifail


CASES = [
    (12, [0.3, 0.5], 1),
    (20, [0.1], 1),
    (16, [0.05, 0.2], 2),
    (24, [1e-3], 2),
    (16, [2e-5], 2),
]


@pytest.mark.parametrize("nx,xis,grid_type", CASES)
def test_kernel_matrices_match_oracle(tk, tko, nx, xis, grid_type):
    xis = np.array(xis)
    tk.matrices.initialize_kernels(nx, xis, grid_type=grid_type)
    tko.initialize_kernels(nx, xis, grid_type, mode=1)
    # QAGS path exactness: same number of bisections and same error flag for every instance
    res, last, ifail = _qags_gpu(tk)
    x_nodes = [tko.grid_nodes(nx, s, grid_type) for s in xis]
    for k, s in enumerate(xis):
        for i in range(nx):
            for c in range(7):
                v, l, e = tko.cst_integral(c, x_nodes[k][i], s)
                assert list(last[k, i, c]) == l and list(ifail[k, i, c]) == e, (k, i, c, last[k, i, c], l)
                got = res[k, i, c]
                tot = (0.0 + got[0]) + got[1]
                tot = tot + got[2]
                assert tot == v or (np.isnan(tot) and np.isnan(v)), (k, i, c, tot, v)
    for nfl in (3, 4, 5):
        for which, name in enumerate(tko.RAW_NAMES):
            got = _raw_gpu(tk, which, nfl)
            ref = tko.raw_kernel(name, nfl)
            close(got, ref)
            assert np.array_equal(got, ref, equal_nan=True), "%s nfl=%d not bit-identical" % (name, nfl)


@pytest.mark.parametrize("variant", [(1, 0), (0, 1), (1, 1)])
def test_ofast_variants_bit_identical(tk, tko, variant):
    """the two float32-division variants of an -Ofast reference build; GPU and oracle take the
    same switches and stay bit-identical in each (the canonical (0, 0) variant is every other test)."""
    nx, xis, grid_type = 24, np.array([0.07, 0.4]), 2
    f90 = tk.f90wrap.matrices.dummy
    try:
        tko.set_variant(0, 0)
        tko.initialize_kernels(nx, xis, grid_type, mode=1)
        canon = tko.raw_kernel(tko.RAW_NAMES[0], 3).copy()
        f90.set_variant(*variant)
        tko.set_variant(*variant)
        tk.matrices.initialize_kernels(nx, xis, grid_type=grid_type)
        tko.initialize_kernels(nx, xis, grid_type, mode=1)
        for which, name in enumerate(tko.RAW_NAMES):
            got = _raw_gpu(tk, which, 4)
            ref = tko.raw_kernel(name, 4)
            assert np.array_equal(got, ref, equal_nan=True), "%s variant %s not bit-identical" % (name, variant)
        # the switch does something: entries move at the 1e-4 level, not the 1e-15 level
        moved = np.abs(tko.raw_kernel(tko.RAW_NAMES[0], 3) - canon).max() / np.abs(canon).max()
        assert 1e-9 < moved < 1e-2, moved
    finally:
        f90.set_variant(0, 0)
        tko.set_variant(0, 0)


@pytest.mark.parametrize("nlo", [False, True])
def test_evolution_matrices_match_oracle(tk, tko, nlo):
    nx, xis, gt = 16, np.array([0.1, 0.4]), 1
    Q2 = tk.matrices.Q2space(4.0, 100.0, 6)          # crosses mb2: nfl 4 -> 5
    tk.matrices.initialize_kernels(nx, xis, grid_type=gt)
    tk.matrices.initialize_evolution_matrices(Q2, nlo=nlo)
    tko.initialize_kernels(nx, xis, gt, mode=1)
    tko.initialize_evolution_matrices(Q2, nlo=nlo)
    for name, kw in (("matrix_VNS", {"ns_type": 1}), ("matrix_VNS", {"ns_type": -1}),
                     ("matrix_ANS", {"ns_type": 1}), ("matrix_ANS", {"ns_type": -1}),
                     ("matrix_VSG", {}), ("matrix_ASG", {})):
        got = getattr(tk.matrices, name)(**kw)
        assert got.is_cuda and got.dtype == torch.float64
        ref = getattr(tko, name)(**kw)
        assert tuple(got.shape) == ref.shape
        close(got.cpu().numpy(), ref)
    # identity at the first Q2, LO plus == minus, A-NS is the swapped V-NS (matevo.f90:263-272, 505-510)
    M = tk.matrices.matrix_VNS().cpu().numpy()
    assert np.array_equal(M[:, :, :, 0], np.broadcast_to(np.eye(nx)[:, :, None], (nx, nx, 2)))
    if not nlo:
        assert torch.equal(tk.matrices.matrix_VNS(1), tk.matrices.matrix_VNS(-1))
    assert torch.equal(tk.matrices.matrix_ANS(1), tk.matrices.matrix_VNS(-1))
    assert torch.equal(tk.matrices.matrix_ANS(-1), tk.matrices.matrix_VNS(1))


def test_kernel_getters_match_oracle(tk, tko):
    nx, xis, gt = 12, np.array([0.2, 0.6]), 1
    tk.matrices.initialize_kernels(nx, xis, grid_type=gt)
    tko.initialize_kernels(nx, xis, gt, mode=1)
    for kind in ("VQQ", "VQG", "VGQ", "VGG", "AQQ", "AQG", "AGQ", "AGG"):
        for nlo in (False, True):
            for nfl in (3, 5):
                nss = (1, -1, 0) if kind.endswith("QQ") else (1,)
                for ns in nss:
                    kw = {"Q2": 7.5, "nfl": nfl, "nlo": nlo}
                    if kind.endswith("QQ"):
                        kw["ns_type"] = ns
                    got = getattr(tk.matrices, "kernel_" + kind)(**kw).cpu().numpy()
                    ref = tko.kernel_matrix(kind, Q2=7.5, nfl=nfl, nlo=nlo, ns_type=ns)
                    close(got, ref, rtol=1e-13)


# --------------------------------------------------------------------------------------------
# application

@pytest.mark.parametrize("S", [1, 3, 4, 5, 8, 9, 16, 23, 40, 64, 65, 200])
def test_apply_matches_einsum(tk, S):
    rng = np.random.default_rng(5)
    m, nxi, nq = 24, 3, 4
    M = rng.standard_normal((m, m, nxi, nq))
    g = rng.standard_normal((m, nxi, S))
    out = tk.matrices.apply_matrix(dev(M), g)
    assert tuple(out.shape) == (m, nxi, S, nq)
    ref = np.einsum('ijkl,jk...->ik...l', M, g)
    close(out.cpu().numpy(), ref, rtol=1e-12, afloor=1e-13)
    # extra trailing dimensions ride along, as in examples/examples.py:140-148
    g2 = rng.standard_normal((m, nxi, 2, 5))
    out2 = tk.matrices.apply_matrix(dev(M), g2)
    close(out2.cpu().numpy(), np.einsum('ijkl,jk...->ik...l', M, g2), rtol=1e-12, afloor=1e-13)
    # no sample axis at all
    g0 = rng.standard_normal((m, nxi))
    out0 = tk.matrices.apply_matrix(dev(M), g0)
    close(out0.cpu().numpy(), np.einsum('ijkl,jk...->ik...l', M, g0), rtol=1e-12, afloor=1e-13)


def test_evolved_gk_gpds_match_oracle(tk, tko):
    """config C1-style: GK input, LO and NLO, identical inputs on both sides (examples/examples.py:37-99)"""
    nx, xi = 20, 0.5
    Q2 = np.array([4.0, 8.0, 12.0, 16.0])
    x = tko.pixelspace(nx)
    umin0 = tko.gk("Hu", x, [xi], [0.0]) + tko.gk("Hu", -x, [xi], [0.0])
    sing0 = tko.H_singlet(x, [xi], [0.0])
    glue0 = tko.gk("Hg", x, [xi], [0.0])
    sg0 = np.concatenate((sing0, glue0), axis=0)
    tk.matrices.initialize_kernels(nx, xi)
    tko.initialize_kernels(nx, [xi], 1, mode=1)
    for nlo in (False, True):
        tk.matrices.initialize_evolution_matrices(Q2, nlo=nlo)
        tko.initialize_evolution_matrices(Q2, nlo=nlo)
        got = tk.matrices.apply_matrix(tk.matrices.matrix_VNS(), umin0).cpu().numpy()
        ref = np.einsum('ijkl,jk...->ik...l', tko.matrix_VNS(), umin0)
        close(got, ref)
        got = tk.matrices.apply_matrix(tk.matrices.matrix_VSG(), sg0).cpu().numpy()
        ref = np.einsum('ijkl,jk...->ik...l', tko.matrix_VSG(), sg0)
        close(got, ref)
