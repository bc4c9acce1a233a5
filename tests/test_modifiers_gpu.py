"""The stand-alone forms of the modifiers for the seams next to TensorProductScatter (nequip_b200/nn/modifiers.py)."""
import pytest
import torch

from nequip_b200 import data as D
from nequip_b200.nn import modifiers as M
from oracle import sh as osh

pytestmark = pytest.mark.gpu


def test_edge_embed_modifier_swaps_and_matches_oracle():
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.sh = M._SHInterface(2)

    net = M.enable_B200EdgeEmbed(Net())
    assert isinstance(net.sh, M.B200SphericalHarmonicEdgeAttrs)
    sysd = D.make_system("water", 4, r_max=5.0, seed=1)
    sysd.pop("_meta")
    out = net.sh(D.to_device(sysd, "cuda"))
    vec = out["edge_vectors"].cpu()
    ref = osh.spherical_harmonics(2, vec, normalize=True).to(torch.float32)
    torch.testing.assert_close(out["edge_attrs"].cpu(), ref, rtol=1e-6, atol=1e-6)


def test_ghost_exchange_module_single_rank_is_identity():
    from nequip_b200 import parallel as P

    sysd = D.make_system("water", 4, r_max=5.0, seed=1)
    sysd.pop("_meta")
    plan = P.make_plans(sysd["edge_index"], torch.zeros(sysd["pos"].shape[0], dtype=torch.long), 1)[0]
    halo = P.HaloExchange(plan, "cuda")
    mod = M.B200GhostExchangeModule(field="node_features", irreps_in={"node_features": "8x0e"})
    x = torch.randn(plan.n_own, 8, device="cuda")
    out = mod({"node_features": x, M.NQB_HALO_KEY: halo}, ghost_included=False)
    assert torch.equal(out["node_features"], x)
    with pytest.raises(RuntimeError):
        mod({"node_features": x}, ghost_included=False)
