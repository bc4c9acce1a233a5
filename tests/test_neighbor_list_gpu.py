"""GPU cell-list neighbour list (nqb_nl_*, nequip_b200/ops.py neighbor_list) vs the host lists of
nequip_b200/data.py (cell list / brute force with images) -- same contract as the reference's backends
(nequip/data/_nl.py:60-152): the edge set, the shifts and the (centre, neighbour) order must be identical."""
import numpy as np
import pytest
import torch

from nequip_b200 import data as D
from nequip_b200 import ops

pytestmark = pytest.mark.gpu


def _check(pos_np, cell_np, r_max, pbc=True):
    ei_ref, sh_ref = D.neighbor_list(pos_np, cell_np, r_max, pbc=pbc)
    out = ops.neighbor_list(torch.from_numpy(pos_np).cuda(), None if cell_np is None else torch.from_numpy(cell_np), pbc, r_max,
                            transpose_perm=True)
    ei, sh = out["edge_index"].cpu().numpy(), out["edge_cell_shift"].cpu().numpy()
    assert ei.shape == ei_ref.shape, (ei.shape, ei_ref.shape)
    np.testing.assert_array_equal(ei, ei_ref)
    np.testing.assert_array_equal(sh, sh_ref)
    N = pos_np.shape[0]
    rp = out["row_ptr"].cpu().numpy()
    np.testing.assert_array_equal(rp, np.concatenate([[0], np.cumsum(np.bincount(ei_ref[0], minlength=N))]))
    tp = out["edge_transpose_perm"].cpu().numpy()
    key = ei[1][tp] * N + ei[0][tp]
    assert np.all(np.diff(key) >= 0)
    return ei.shape[1]


@pytest.mark.parametrize("kind,n_side", [("li3po4", 12), ("water", 10), ("asi", 16)])
def test_matches_host_cell_list(kind, n_side):
    pr = D.PRESETS[kind]
    pos, cell = D.jittered_lattice(n_side, pr["density"], seed=3)
    pos = pos + np.array([3.7, -11.2, 0.4])  # atoms outside the home cell: base shifts are exercised
    E = _check(pos, cell, 5.0)
    assert E > 0


def test_small_cells_need_several_images():
    rng = np.random.default_rng(0)
    for L in (3.0, 6.5, 9.0):  # r_max = 5 > L/2: the same neighbour appears under several shifts
        cell = np.diag([L, L * 1.1, L * 0.9])
        pos = rng.uniform(0, 1, (11, 3)) @ cell
        _check(pos, cell, 5.0)


def test_non_periodic_and_empty():
    rng = np.random.default_rng(1)
    pos = rng.uniform(0, 14, (200, 3))
    _check(pos, None, 5.0, pbc=False)
    out = ops.neighbor_list(torch.tensor([[0.0, 0, 0], [100.0, 0, 0]], dtype=torch.float64).cuda(), None, False, 5.0)
    assert out["edge_index"].shape == (2, 0) and out["row_ptr"].tolist() == [0, 0, 0]


def test_triclinic_cell_against_bruteforce():
    rng = np.random.default_rng(2)
    cell = np.array([[11.0, 0.0, 0.0], [3.0, 10.0, 0.0], [-2.0, 1.5, 12.0]])
    pos = rng.uniform(0, 1, (150, 3)) @ cell
    out = ops.neighbor_list(torch.from_numpy(pos).cuda(), torch.from_numpy(cell), True, 4.0)
    ei, sh = out["edge_index"].cpu().numpy(), out["edge_cell_shift"].cpu().numpy()
    # brute force over images
    ref = set()
    for a in range(-2, 3):
        for b in range(-2, 3):
            for c in range(-2, 3):
                s = np.array([a, b, c], dtype=np.float64)
                d = pos[None, :, :] + s @ cell - pos[:, None, :]
                ok = (d * d).sum(-1) < 16.0
                if a == 0 and b == 0 and c == 0:
                    ok &= ~np.eye(pos.shape[0], dtype=bool)
                for i, j in zip(*np.nonzero(ok)):
                    ref.add((int(i), int(j), a, b, c))
    got = {(int(i), int(j), int(s[0]), int(s[1]), int(s[2])) for i, j, s in zip(ei[0], ei[1], sh)}
    assert got == ref
    assert np.all(np.diff(ei[0] * pos.shape[0] + ei[1]) >= 0)


def test_model_runs_on_the_device_list():
    from nequip_b200.nn.model import NequIPEnergyModel

    sysd = D.make_system("water", 6, r_max=5.0, seed=0)
    meta = sysd.pop("_meta")
    model = NequIPEnergyModel(r_max=5.0, type_names=meta["type_names"], l_max=2, num_layers=3, num_features=32,
                              avg_num_neighbors=meta["avg_num_neighbors"]).cuda()
    dev = D.to_device(sysd, "cuda")
    ref = model(dev)
    nl = ops.neighbor_list(dev["pos"], sysd["cell"], True, 5.0)
    d2 = dict(dev)
    d2["edge_index"], d2["edge_cell_shift"] = nl["edge_index"], nl["edge_cell_shift"]
    out = model(d2)
    assert torch.equal(out["forces"], ref["forces"]) or float((out["forces"] - ref["forces"]).abs().max()) <= 1e-6 * float(ref["forces"].abs().max())
