"""CPU: pin the oracle by mathematical identities (no e3nn on disk -> "parity unpinned" in the
strict sense, see oracle/__init__.py).  These are the checks SURVEY.md section 8c lists:
closed-form values, Wigner-D equivariance of w3j and SH, the SH recurrence that ties SH signs to
w3j signs, TP equivariance (pins path ordering / normalisation consistency), the gate constants."""
import math

import numpy as np
import pytest
import torch

from oracle import irreps as I
from oracle import sh as osh
from oracle import tp as otp
from oracle import wigner

ANGLES = (0.3, 1.1, -0.7)

NNZ = {
    (1, 1, 1): 6, (1, 1, 2): 11, (1, 2, 2): 16, (2, 2, 2): 25, (1, 2, 3): 21, (1, 3, 3): 26,
    (2, 2, 3): 28, (2, 3, 3): 41, (3, 3, 3): 42, (0, 2, 2): 5, (2, 2, 0): 5, (3, 0, 3): 7,
}


def test_w3j_known_values():
    assert wigner.wigner_3j(0, 0, 0)[0, 0, 0] == pytest.approx(1.0)
    np.testing.assert_allclose(wigner.wigner_3j(1, 1, 0)[:, :, 0], np.eye(3) / math.sqrt(3), atol=1e-15)
    C = wigner.wigner_3j(1, 1, 1)
    eps = np.zeros((3, 3, 3))
    for (i, j, k), s in {(0, 1, 2): 1, (1, 2, 0): 1, (2, 0, 1): 1, (0, 2, 1): -1, (2, 1, 0): -1, (1, 0, 2): -1}.items():
        eps[i, j, k] = s
    np.testing.assert_allclose(C, eps / math.sqrt(6), atol=1e-15)
    for l in range(4):
        np.testing.assert_allclose(wigner.wigner_3j(0, l, l)[0], np.eye(2 * l + 1) / math.sqrt(2 * l + 1), atol=1e-15)
        np.testing.assert_allclose(wigner.wigner_3j(l, 0, l)[:, 0], np.eye(2 * l + 1) / math.sqrt(2 * l + 1), atol=1e-15)


@pytest.mark.parametrize("ls", sorted(NNZ))
def test_w3j_norm_nnz_equivariance(ls):
    C = wigner.wigner_3j(*ls)
    assert np.linalg.norm(C) == pytest.approx(1.0, abs=1e-14)
    assert int((np.abs(C) > 1e-12).sum()) == NNZ[ls]
    Ds = [wigner.wigner_D(l, *ANGLES) for l in ls]
    C2 = np.einsum("ijk,il,jm,kn->lmn", C, *Ds)
    np.testing.assert_allclose(C2, C, atol=5e-15)


def test_wigner_D1_is_cartesian_rotation():
    R = wigner.wigner_D(1, *ANGLES)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)
    assert np.linalg.det(R) == pytest.approx(1.0)
    # y is the polar axis: rotation by alpha about y leaves e_y fixed
    Ry = wigner.wigner_D(1, 0.4, 0.0, 0.0)
    np.testing.assert_allclose(Ry @ np.array([0, 1.0, 0]), [0, 1.0, 0], atol=1e-15)


@pytest.mark.parametrize("lmax", [1, 2, 3])
def test_sh_closed_form_vs_recurrence_and_equivariance(lmax):
    g = torch.Generator().manual_seed(0)
    v = torch.randn(64, 3, generator=g, dtype=torch.float64)
    Yc = osh.sh_closed_form(lmax, v)
    Yr = osh.sh_recurrence(lmax, v)
    torch.testing.assert_close(Yc, Yr, atol=1e-13, rtol=0)
    R = torch.from_numpy(wigner.wigner_D(1, *ANGLES))
    Yrot = osh.sh_closed_form(lmax, v @ R.T)
    off = 0
    for l in range(lmax + 1):
        D = torch.from_numpy(wigner.wigner_D(l, *ANGLES))
        blk = slice(off, off + 2 * l + 1)
        torch.testing.assert_close(Yrot[:, blk], Yc[:, blk] @ D.T, atol=1e-13, rtol=0)
        torch.testing.assert_close((Yc[:, blk] ** 2).sum(-1), torch.full((64,), 2.0 * l + 1, dtype=torch.float64))
        off += 2 * l + 1
    # scale invariance (normalize=True)
    torch.testing.assert_close(osh.sh_closed_form(lmax, 3.7 * v), Yc, atol=1e-13, rtol=0)


def _block_diag_D(irreps, angles):
    mats = []
    for mul, (l, p) in irreps:
        D = wigner.wigner_D(l, *angles)
        for _ in range(mul):
            mats.append(D)
    n = sum(m.shape[0] for m in mats)
    out = np.zeros((n, n))
    o = 0
    for m in mats:
        out[o: o + m.shape[0], o: o + m.shape[0]] = m
        o += m.shape[0]
    return torch.from_numpy(out)


def test_tp_scatter_oracle_is_equivariant():
    fin = I.parse("3x0e+2x1o+2x1e+2x2e")
    fe = I.spherical_harmonics(2)
    fout = I.parse("4x0e+4x0o+4x1e+4x1o+4x2e+4x2o")
    mid, ins = I.build_tp_instructions(fin, fe, fout)
    g = torch.Generator().manual_seed(1)
    N, E = 5, 17
    x = torch.randn(N, I.dim(fin), generator=g, dtype=torch.float64)
    vec = torch.randn(E, 3, generator=g, dtype=torch.float64)
    w = torch.randn(E, otp.weight_numel(fin, fe, ins), generator=g, dtype=torch.float64)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, N, (E,), generator=g)
    R = torch.from_numpy(wigner.wigner_D(1, *ANGLES))
    Din, Dout = _block_diag_D(fin, ANGLES), _block_diag_D(mid, ANGLES)
    out = otp.tp_scatter(x, osh.spherical_harmonics(2, vec), w, dst, src, fin, fe, mid, ins)
    out_rot = otp.tp_scatter(x @ Din.T, osh.spherical_harmonics(2, vec @ R.T), w, dst, src, fin, fe, mid, ins)
    torch.testing.assert_close(out_rot, out @ Dout.T, atol=1e-12, rtol=0)
    # parity: inversion flips odd irreps of input and output consistently
    sgn_in = torch.cat([torch.full((m * (2 * l + 1),), float(p)) for m, (l, p) in fin]).double()
    sgn_out = torch.cat([torch.full((m * (2 * l + 1),), float(p)) for m, (l, p) in mid]).double()
    out_inv = otp.tp_scatter(x * sgn_in, osh.spherical_harmonics(2, -vec), w, dst, src, fin, fe, mid, ins)
    torch.testing.assert_close(out_inv, out * sgn_out, atol=1e-12, rtol=0)


def test_tp_chunked_equals_unchunked():
    fin, fe = I.parse("4x0e+4x1o"), I.spherical_harmonics(1)
    mid, ins = I.build_tp_instructions(fin, fe, I.parse("4x0e+4x1o+4x1e"))
    g = torch.Generator().manual_seed(2)
    x = torch.randn(6, I.dim(fin), generator=g, dtype=torch.float64)
    y = torch.randn(40, 4, generator=g, dtype=torch.float64)
    w = torch.randn(40, otp.weight_numel(fin, fe, ins), generator=g, dtype=torch.float64)
    src, dst = torch.randint(0, 6, (40,), generator=g), torch.randint(0, 6, (40,), generator=g)
    a = otp.tp_scatter(x, y, w, dst, src, fin, fe, mid, ins)
    b = otp.tp_scatter(x, y, w, dst, src, fin, fe, mid, ins, chunk=7)
    torch.testing.assert_close(a, b, atol=1e-13, rtol=0)


def test_path_normalisation():
    """InteractionBlock gives every instruction its own output -> coefficient sqrt(2 l3 + 1)."""
    fin, fe = I.parse("2x0e+2x1o+2x2e"), I.spherical_harmonics(2)
    mid, ins = I.build_tp_instructions(fin, fe, I.parse("2x0e+2x1o+2x1e+2x2e+2x2o"))
    for (a, b, c, _, _), coef in zip(ins, otp.path_coefficients(fin, fe, mid, ins)):
        assert coef == pytest.approx(math.sqrt(2 * mid[c][1][0] + 1))


def test_normalize2mom_constants():
    from oracle import model as om

    z = torch.randn(1_000_000, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    assert torch.nn.functional.silu(z).pow(2).mean().pow(-0.5).item() == pytest.approx(om.C_SILU, rel=1e-12)
    assert torch.tanh(z).pow(2).mean().pow(-0.5).item() == pytest.approx(om.C_TANH, rel=1e-12)


def test_polynomial_cutoff_properties():
    from oracle import model as om

    x = torch.tensor([0.0, 0.5, 0.999999, 1.0, 1.2], dtype=torch.float64, requires_grad=True)
    f = om.polynomial_cutoff(x, 6.0)
    assert f[0].item() == 1.0 and f[3].item() == 0.0 and f[4].item() == 0.0 and abs(f[2].item()) < 1e-13
    (g,) = torch.autograd.grad(f.sum(), x)
    assert abs(g[2].item()) < 1e-9


def test_oracle_model_finite_difference_forces():
    """model_tests_basic.py:631-672 restated on the oracle (float64)."""
    from nequip_b200 import data as D
    from nequip_b200.nn.model import NequIPEnergyModel
    from oracle import model as om

    sysd = D.make_system("water", 3, seed=2)
    meta = sysd.pop("_meta")
    m = NequIPEnergyModel(r_max=5.0, type_names=meta["type_names"], avg_num_neighbors=meta["avg_num_neighbors"],
                          l_max=2, num_layers=3, num_features=4, radial_mlp_width=8, model_dtype=torch.float64)
    e, ea, f = om.energy_and_forces(m.state_dict(), m.config, sysd, torch.float64)
    assert float(f.sum(0).abs().max()) < 1e-12  # translation invariance
    eps = 1e-5
    for (i, c) in [(0, 0), (7, 2)]:
        es = []
        for s in (1, -1):
            d = dict(sysd)
            p = sysd["pos"].clone()
            p[i, c] += s * eps
            d["pos"] = p
            es.append(om.energy(m.state_dict(), m.config, d, torch.float64)[0].item())
        assert -(es[0] - es[1]) / (2 * eps) == pytest.approx(f[i, c].item(), rel=1e-6, abs=1e-9)


def test_su2_clebsch_gordan_against_sympy():
    """Independent published implementation (sympy.physics.quantum.cg.CG, Condon-Shortley convention): the oracle's
    Racah-formula SU(2) coefficients, which the real Wigner-3j tensors are built from, agree for every l <= 3 triple."""
    sympy = pytest.importorskip("sympy")
    from sympy.physics.quantum.cg import CG

    for l1 in range(4):
        for l2 in range(4):
            for l3 in range(abs(l1 - l2), min(3, l1 + l2) + 1):
                ours = wigner.su2_cg(l1, l2, l3)
                for m1 in range(-l1, l1 + 1):
                    for m2 in range(-l2, l2 + 1):
                        m3 = m1 + m2
                        if abs(m3) > l3:
                            continue
                        ref = float(CG(l1, m1, l2, m2, l3, m3).doit())
                        assert abs(float(ours[l1 + m1, l2 + m2, l3 + m3]) - ref) < 1e-12, (l1, l2, l3, m1, m2)


def test_real_sh_against_scipy_complex_harmonics():
    """Independent implementation (scipy.special sph_harm): the oracle's real harmonics are, up to the e3nn
    convention's fixed real-basis change (m < 0 sine-like, m > 0 cosine-like, axis order y, z, x) and the
    'component' normalisation sqrt(4 pi), the standard real spherical harmonics."""
    scipy_special = pytest.importorskip("scipy.special")
    sph = getattr(scipy_special, "sph_harm_y", None)
    g = torch.Generator().manual_seed(0)
    v = torch.randn(50, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=1, keepdim=True)
    Y = osh.spherical_harmonics(3, v, normalize=True).numpy()
    # e3nn convention: the "z" axis of the standard harmonics is the y component, (x, y) -> (z, x)
    x, y, z = v[:, 2].numpy(), v[:, 0].numpy(), v[:, 1].numpy()
    theta, phi = np.arccos(np.clip(z, -1, 1)), np.arctan2(y, x)
    off = 0
    for l in range(4):
        for m in range(-l, l + 1):
            if sph is not None:
                c = sph(l, abs(m), theta, phi)
            else:
                c = scipy_special.sph_harm(abs(m), l, phi, theta)
            if m == 0:
                std = c.real
            elif m > 0:
                std = np.sqrt(2) * (-1) ** m * c.real
            else:
                std = np.sqrt(2) * (-1) ** m * c.imag
            ours = Y[:, off + l + m] / np.sqrt(4 * np.pi)
            # sign conventions per (l, m) are fixed constants: compare up to that sign, and require it to be uniform
            s = np.sign(np.sum(ours * std))
            assert s != 0
            np.testing.assert_allclose(ours, s * std, atol=1e-12, err_msg=f"l={l} m={m}")
        off += 2 * l + 1
