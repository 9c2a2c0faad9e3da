"""CPU tests of the host-side layer: the C-ABI library loads and exports every declared symbol,
the numpy/torch-free helpers behave like the reference's (tiktaalik/matrices.py:95-130,
tiktaalik/evolution.py:435-548), the xi partition, the GK model, and the no-fallback rule."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tiktaalik_b200 import _lib
    L = _lib.lib()                                  # loads on a machine without a GPU
    hdr = open(os.path.join(ROOT, "include", "tiktaalik_b200.h")).read()
    declared = set(re.findall(r"\b(tkb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name)
    assert L.tkb_version() >= 100
    assert isinstance(L.tkb_last_error(), bytes)


def test_host_only_entry_points_match_oracle(tko):
    import tiktaalik_b200 as tk
    for nx, xi, gt in ((20, 0.5, 1), (80, 0.1, 2), (16, 2e-5, 2)):
        np.testing.assert_allclose(tk.matrices.pixelspace(nx, xi, gt), tko.pixelspace(nx, xi, gt), rtol=4e-16)
    q2 = np.array([1.0, tk.pars.mc2, 4.0, 10.0, tk.pars.mb2, 100.0, 91.1876 ** 2])
    np.testing.assert_allclose(tk.qcd.alphaQCD(q2), tko.alphaQCD(q2), rtol=1e-14)
    assert tk.qcd.alphaQCD(91.1876 ** 2)[0] == 0.118


def test_compute_entry_points_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import tiktaalik_b200 as tk
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tk.matrices.initialize_kernels(12, 0.5)
    with pytest.raises(RuntimeError):
        tk.matrices.apply_matrix(np.zeros((4, 4, 1, 2)), np.zeros((4, 1)))
    with pytest.raises(RuntimeError):
        tk.Evolver(nx=12, xi=0.5)


def test_product_never_touches_the_oracle():
    """nothing under tiktaalik_b200/ may include, link, import or execute anything from oracle/"""
    pkg = os.path.join(ROOT, "tiktaalik_b200")
    bad = re.compile(r"#include\s*[\"<][^\">]*oracle|\bimport\s+oracle|\bfrom\s+oracle|libtko|\btko_|\btko\b|oracle/")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_q2space_passes_get_nfl():
    import tiktaalik_b200 as tk
    m = tk.matrices
    q = m.Q2space(4.0, 10.0, 20)
    assert q.shape == (20,) and np.allclose(q, np.geomspace(4.0, 10.0, 20))
    q = m.Q2space(1.0, 100.0, 8)
    assert q.shape == (8,) and tk.pars.mc2 in q and tk.pars.mb2 in q and np.all(np.diff(q) > 0)
    assert m.passes(q, tk.pars.mc2) and not m.passes(np.array([4.0, 10.0]), tk.pars.mb2)
    assert not m.passes(np.array([tk.pars.mc2, 10.0]), tk.pars.mc2)
    assert [m.get_nfl(v) for v in (1.0, tk.pars.mc2, 4.0, tk.pars.mb2, 100.0)] == [3, 4, 4, 5, 5]
    assert tk.pars.mc2 == 1.6641000000000001 and tk.pars.mb2 == 17.472399999999997


def test_basis_helpers_on_cpu_tensors():
    from tiktaalik_b200 import evolution as ev
    for nfl in (3, 4, 5):
        T = ev.Tpls_matrix(nfl)
        assert T.shape == (nfl, nfl) and np.all(T[0] == 1) and T[nfl - 1, nfl - 1] == -(nfl - 1)
        assert abs(np.linalg.det(T)) > 0
    g = np.arange(24.0).reshape(2, 3, 4)
    p = ev.add_pad(g, 3)
    assert p.shape == (5, 3, 4) and np.all(p[2:] == 0) and np.array_equal(p[:2], g)
    t = ev.add_pad(torch.ones(2, 3), 1)
    assert tuple(t.shape) == (3, 3) and float(t[2].sum()) == 0
    assert ev.swap_AV('A') == 'V' and ev.swap_AV('V') == 'A'
    with pytest.raises(ValueError):
        ev.swap_AV('X')


def test_xi_partition_and_clamp():
    from tiktaalik_b200 import shard
    counts, offs = shard.xi_partition(10, 4, "blocks")
    assert counts == [3, 3, 2, 2] and offs == [0, 3, 6, 8]
    counts, offs = shard.xi_partition(10, 4)                                   # strided (default): rank r holds r, r + 4, ...
    assert counts == [3, 3, 2, 2] and offs == [0, 1, 2, 3]
    xi = np.array([1e-7, 0.1, 0.5, 0.9999999])
    c = shard.clamp_xi(xi)
    assert c[0] == 1e-6 and c[-1] == 0.999999 and xi[0] == 1e-7               # matevo.f90:688-689, input untouched
    parts = [shard.local_xi(xi, r, 2, "blocks")[0] for r in range(2)]
    assert np.array_equal(np.concatenate(parts), c)
    parts = [shard.local_xi(xi, r, 2)[0] for r in range(2)]
    assert np.array_equal(parts[0], c[0::2]) and np.array_equal(parts[1], c[1::2])


def test_gk_model_matches_oracle(tko):
    import tiktaalik_b200 as tk
    x = tk.matrices.pixelspace(40)
    t = np.array([0.0, -0.3])
    for name in ("Hu", "Hd", "Hs", "Hg", "Hu_val", "Hd_val", "Hu_sea", "Hd_sea"):
        for xi in (0.5, 0.2):
            a = getattr(tk.model.host, name)(x, xi, t)
            b = tko.gk(name, x, [xi], t)
            assert a.shape == (40, 1, 2)
            assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    a = tk.model.host.H_singlet(x, np.array([0.3, 0.5]), 0.0)
    assert a.shape == (40, 2, 1)
    assert np.abs(a - tko.H_singlet(x, [0.3, 0.5], [0.0])).max() <= 1e-9 * np.abs(a).max()
    assert np.allclose(a, -a[::-1])                                            # the singlet is odd in x


def test_matrix_cache_round_trip(tmp_path):
    from tiktaalik_b200 import cache
    rng = np.random.default_rng(3)
    nx, nxi, nq = 6, 5, 3
    mats = {"VNS_plus": torch.as_tensor(rng.standard_normal((nx, nx, nxi, nq))),
            "VSG": torch.as_tensor(rng.standard_normal((2 * nx, 2 * nx, nxi, nq)))}
    xi = np.linspace(0.1, 0.9, nxi)
    meta = cache.save_matrices(str(tmp_path / "c"), mats, nx, xi, [4.0, 6.0, 10.0], nlo=True, grid_type=2)
    assert meta["types"] == ["VNS_plus", "VSG"]
    raw = np.load(str(tmp_path / "c" / "VSG.npy"))
    assert raw.shape == (nxi, nq, 2 * nx, 2 * nx) and raw.flags["C_CONTIGUOUS"]
    assert np.array_equal(raw.transpose(2, 3, 0, 1), mats["VSG"].numpy())           # plain .npy, reference index order
    meta2, back = cache.load_matrices(str(tmp_path / "c"), device=torch.device("cpu"))
    assert meta2["nx"] == nx and meta2["nlo"] is True and meta2["grid_type"] == 2
    for k in mats:
        assert tuple(back[k].shape) == tuple(mats[k].shape) and torch.equal(back[k], mats[k])
    meta3, part = cache.load_matrices(str(tmp_path / "c"), device=torch.device("cpu"), xi_slice=slice(1, 3))
    assert torch.equal(part["VNS_plus"], mats["VNS_plus"][:, :, 1:3, :]) and len(meta3["xi"]) == 2
    with pytest.raises(ValueError):
        cache.save_matrices(str(tmp_path / "d"), {"bogus": mats["VSG"]}, nx, xi, [4.0, 6.0, 10.0], nlo=True)


def test_integration_document_names_every_c_abi_entry():
    """INTEGRATION.md is the binding guide: every function include/tiktaalik_b200.h declares appears in it"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "tiktaalik_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(tkb_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 25
    assert [n for n in names if n not in doc] == []


def _table_literals(path, name):
    """the hexadecimal literals of one table of a generated header, in order"""
    txt = open(path).read()
    body = txt[txt.index(name):]
    body = body[body.index("{"):body.index("};")]
    return re.findall(r"-?0x[0-9a-fA-F.]+(?:p[+-]?\d+)?", body)


def test_log_pow_tables_are_the_same_on_both_sides():
    """oracle/tko_pow_tables.h and tiktaalik_b200/csrc/tkb_pow_tables.cuh come from one generator (tools/make_pow_tables.py);
    bit-identical log / pow on host and device needs bit-identical tables"""
    c = os.path.join(ROOT, "oracle", "tko_pow_tables.h")
    cu = os.path.join(ROOT, "tiktaalik_b200", "csrc", "tkb_pow_tables.cuh")
    for cname, cuname, n in (("TKO_POW_LOG", "c_pow_log", 93 * 3), ("TKO_POW_EXP", "c_pow_exp", 128 * 2),
                             ("TKO_LOG_TAB", "g_log_tab", 129 * 3)):
        a, b = _table_literals(c, cname), _table_literals(cu, cuname)
        assert len(a) == n and a == b, cname
    inv = [re.search(r"INV_LN2_128 = (\S+);", open(p).read()).group(1) for p in (c, cu)]
    assert inv[0] == inv[1]


def test_log_pow_tables_regenerate_identically(tmp_path):
    """the committed tables are what tools/make_pow_tables.py writes (needs mpmath)"""
    pytest.importorskip("mpmath")
    import importlib.util
    import shutil
    spec = importlib.util.spec_from_file_location("make_pow_tables", os.path.join(ROOT, "tools", "make_pow_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for sub in ("oracle", os.path.join("tiktaalik_b200", "csrc")):
        os.makedirs(tmp_path / sub)
    mod.ROOT = str(tmp_path)
    mod.main()
    for rel in (os.path.join("oracle", "tko_pow_tables.h"), os.path.join("tiktaalik_b200", "csrc", "tkb_pow_tables.cuh")):
        assert open(tmp_path / rel).read() == open(os.path.join(ROOT, rel)).read(), rel
    shutil.rmtree(tmp_path, ignore_errors=True)
