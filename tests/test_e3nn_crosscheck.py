"""Opportunistic pin of the oracle to the real thing: runs only where ``import e3nn`` succeeds (it does not in
the build container nor on the GPU boxes -- SURVEY F4 -- so this file is normally skipped).  e3nn (pinned
``>=0.6.0,<0.7.0`` by the reference, pyproject.toml:21) is where the arithmetic of the hot path lives; the
reference call sites are nequip/nn/_tp_scatter_base.py:24-31 (TensorProduct), nequip/nn/embedding/_edge.py:187-189
(SphericalHarmonics) and nequip/nn/interaction_block.py:82-87 (Linear)."""
import pytest
import torch

e3nn = pytest.importorskip("e3nn")
from e3nn import o3  # noqa: E402

from oracle import irreps as I  # noqa: E402
from oracle import sh as osh  # noqa: E402
from oracle import tp as otp  # noqa: E402
from oracle import wigner as ow  # noqa: E402


@pytest.mark.parametrize("l1,l2,l3", [(a, b, c) for a in range(4) for b in range(4) for c in range(abs(a - b), min(3, a + b) + 1)])
def test_wigner_3j(l1, l2, l3):
    ref = o3.wigner_3j(l1, l2, l3, dtype=torch.float64)
    torch.testing.assert_close(torch.from_numpy(ow.wigner_3j(l1, l2, l3)), ref, rtol=0, atol=1e-12)


def test_spherical_harmonics_component_normalised():
    g = torch.Generator().manual_seed(0)
    vec = torch.randn(64, 3, generator=g, dtype=torch.float64)
    ref = o3.spherical_harmonics(list(range(4)), vec, normalize=True, normalization="component")
    torch.testing.assert_close(osh.spherical_harmonics(3, vec, normalize=True), ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("fin,fe,fout", [("4x0e+4x1o", "1x0e+1x1o", "4x0e+4x1o+4x1e"),
                                          ("8x0e+8x1o+8x2e", "1x0e+1x1o+1x2e", "8x0e+8x1o+8x2e"),
                                          ("4x0e+4x0o+4x1o+4x1e+4x2e+4x2o", "1x0e+1x1o+1x2e+1x3o", "4x0e+4x1o+4x2e+4x3o")])
def test_tensor_product_uvu(fin, fe, fout):
    mid, ins = I.build_tp_instructions(I.parse(fin), I.parse(fe), I.parse(fout))
    tp = o3.TensorProduct(o3.Irreps(fin), o3.Irreps(fe), o3.Irreps(I.fmt(mid)), ins, shared_weights=False,
                          internal_weights=False).to(torch.float64)
    g = torch.Generator().manual_seed(1)
    E = 15
    x = torch.randn(E, tp.irreps_in1.dim, generator=g, dtype=torch.float64)
    y = torch.randn(E, tp.irreps_in2.dim, generator=g, dtype=torch.float64)
    w = torch.randn(E, tp.weight_numel, generator=g, dtype=torch.float64)
    assert tp.weight_numel == otp.weight_numel(I.parse(fin), I.parse(fe), ins)
    ref = tp(x, y, w)
    got = otp.tensor_product_uvu(x, y, w, I.parse(fin), I.parse(fe), mid, ins)
    torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-12)
