"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).

CPU part: the oracle and the product's own host code (nequip_b200.cg / irreps) reproduce the frozen
numbers.  GPU part: the CUDA kernels reproduce them through the reference-facing module.  The
fixtures do not pin the oracle to the reference (e3nn is absent; "parity unpinned") --
they make drift visible."""
import os

import numpy as np
import pytest
import torch

from nequip_b200 import cg
from nequip_b200.irreps import Irreps, build_tp_instructions
from oracle import irreps as I
from oracle import sh as osh
from oracle import tp as otp
from oracle import wigner as ow

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_w3j_fixture_matches_oracle_and_product():
    z = _load("w3j_lmax3.npz")
    assert len(z.files) > 30
    for key in z.files:
        l1, l2, l3 = (int(c) for c in key.split("_")[1])
        np.testing.assert_allclose(ow.wigner_3j(l1, l2, l3), z[key], rtol=0, atol=1e-14)
        np.testing.assert_allclose(cg.real_w3j(l1, l2, l3), z[key], rtol=0, atol=1e-13)
        assert abs(float((z[key] ** 2).sum()) - 1.0) < 1e-12


def test_sh_fixture_matches_oracle():
    z = _load("sh_lmax3.npz")
    y = osh.spherical_harmonics(3, torch.from_numpy(z["vec"]), normalize=True).numpy()
    np.testing.assert_allclose(y, z["y"], rtol=0, atol=1e-13)
    # component normalisation: sum_m Y_lm^2 = 2l+1
    for l, sl in enumerate([slice(0, 1), slice(1, 4), slice(4, 9), slice(9, 16)]):
        np.testing.assert_allclose((z["y"][:, sl] ** 2).sum(1), 2 * l + 1, rtol=1e-12)


def _cases(z):
    n = len([k for k in z.files if k.endswith("_out")])
    for ci in range(n):
        fin, fe, fout, mid = (str(s) for s in z[f"c{ci}_irreps"])
        t = {k: torch.from_numpy(z[f"c{ci}_{k}"]) for k in ("x", "y", "w", "dst", "src", "out", "go", "gx", "gy", "gw")}
        yield ci, fin, fe, fout, mid, z[f"c{ci}_instructions"], t


def test_tp_scatter_fixture_matches_oracle_and_instruction_builder():
    z = _load("tp_scatter_grid.npz")
    for ci, fin, fe, fout, mid, ins_np, t in _cases(z):
        # the product's instruction builder gives the fixture's path table
        pmid, pins = build_tp_instructions(fin, fe, fout)
        assert str(pmid) == mid and [list(i[:3]) for i in pins] == ins_np.tolist()
        omid, oins = I.build_tp_instructions(I.parse(fin), I.parse(fe), I.parse(fout))
        x, y, w = (t[k].clone().requires_grad_(True) for k in ("x", "y", "w"))
        o = otp.tp_scatter(x, y, w, t["dst"], t["src"], I.parse(fin), I.parse(fe), omid, oins)
        gx, gy, gw = torch.autograd.grad([o], [x, y, w], [t["go"]])
        for got, key in ((o, "out"), (gx, "gx"), (gy, "gy"), (gw, "gw")):
            torch.testing.assert_close(got.detach(), t[key], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_tp_scatter_kernels_match_fixture(dtype, tol):
    from nequip_b200.nn import B200TensorProductScatter

    z = _load("tp_scatter_grid.npz")
    for ci, fin, fe, fout, mid, ins_np, t in _cases(z):
        ins = [(int(a), int(b), int(c), "uvu", True) for a, b, c in ins_np]
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            mod = B200TensorProductScatter(Irreps(fin), Irreps(fe), Irreps(mid), ins)
        finally:
            torch.set_default_dtype(prev)
        x, y, w = (t[k].to("cuda", dtype).requires_grad_(True) for k in ("x", "y", "w"))
        o = mod(x, y, w, t["dst"].cuda(), t["src"].cuda())
        gx, gy, gw = torch.autograd.grad([o], [x, y, w], [t["go"].to("cuda", dtype)])
        for got, key in ((o, "out"), (gx, "gx"), (gy, "gy"), (gw, "gw")):
            ref = t[key]
            # the reference test's criterion (atol = rtol = 1e-5 / 1e-10), scaled to the data magnitude
            torch.testing.assert_close(got.detach().cpu().double(), ref, rtol=tol, atol=tol * float(ref.abs().max()))


@pytest.mark.gpu
def test_sh_kernel_matches_fixture():
    from nequip_b200 import ops

    z = _load("sh_lmax3.npz")
    y = ops.spherical_harmonics(torch.from_numpy(z["vec"]).cuda(), 3, out_dtype=torch.float64)
    np.testing.assert_allclose(y.cpu().numpy(), z["y"], rtol=0, atol=1e-12)
