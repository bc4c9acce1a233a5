"""CUDA-graph replay of the energy+forces step (nequip_b200/graph.py) must reproduce the eager step
bit for bit on the frame it was captured with and on other frames of the same shape."""
import pytest
import torch

from nequip_b200 import data as D
from nequip_b200.graph import GraphedEnergyForces
from nequip_b200.nn.model import NequIPEnergyModel

pytestmark = pytest.mark.gpu


def _build(n_side=6):
    sysd = D.make_system("li3po4", n_side, r_max=5.0, seed=0)
    meta = sysd.pop("_meta")
    model = NequIPEnergyModel(r_max=5.0, type_names=meta["type_names"], parity=True,
                              avg_num_neighbors=meta["avg_num_neighbors"], l_max=2, num_layers=4, num_features=64,
                              radial_mlp_depth=1, radial_mlp_width=128).cuda()
    for p in model.parameters():
        p.requires_grad_(False)
    return model, sysd


def test_graph_replay_matches_eager():
    model, sysd = _build()
    dev = D.to_device(sysd, "cuda")
    eager = model(dev)
    g = GraphedEnergyForces(model, dev)
    out = g(dev)
    g.check_sorted()
    # same kernels, same order; only atomics (red.global.add) may reorder
    torch.testing.assert_close(out["total_energy"], eager["total_energy"], rtol=1e-12, atol=1e-9 * abs(float(eager["total_energy"])))
    fs = float(eager["forces"].abs().max())
    assert float((out["forces"] - eager["forces"]).abs().max()) <= 2e-6 * fs
    assert g.launches_per_replay > 20

    # another frame of the same shape: same edge list, displaced atoms
    gen = torch.Generator().manual_seed(3)
    sys2 = dict(sysd)
    sys2["pos"] = sysd["pos"] + 0.05 * torch.randn(sysd["pos"].shape, generator=gen, dtype=sysd["pos"].dtype)
    dev2 = D.to_device(sys2, "cuda")
    eager2 = model(dev2)
    e2_ref, f2_ref = eager2["total_energy"].clone(), eager2["forces"].clone()
    host2 = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sys2.items()}
    out2 = g(host2)  # host -> static buffers -> replay
    torch.cuda.synchronize()
    assert float((out2["forces"] - f2_ref).abs().max()) <= 2e-6 * float(f2_ref.abs().max())
    assert abs(float(out2["total_energy"]) - float(e2_ref)) <= 1e-9 * abs(float(e2_ref)) + 1e-9
    assert float((f2_ref - eager["forces"]).abs().max()) > 1e-3 * fs  # the frame really changed


def test_graph_rejects_other_shapes_and_unsorted_edges():
    model, sysd = _build(n_side=5)
    dev = D.to_device(sysd, "cuda")
    g = GraphedEnergyForces(model, dev, warmup=1)
    bad = dict(dev)
    bad["pos"] = dev["pos"][:-1]
    assert not g.matches(bad)
    with pytest.raises(ValueError):
        g.load(bad)
    # reversed edge order: not grouped by destination -> flagged after the replay
    rev = dict(dev)
    rev["edge_index"] = dev["edge_index"].flip(1).contiguous()
    rev["edge_cell_shift"] = dev["edge_cell_shift"].flip(0).contiguous()
    g(rev)
    with pytest.raises(RuntimeError):
        g.check_sorted()
