"""The behavioural tests the reference holds for energy models (nequip/utils/unittests/model_tests_basic.py --
SURVEY.md 8c (3)), re-stated against the CPU oracle: they pin the oracle (the checker of every GPU parity test) to the
properties the reference itself demands of this path.  float64 model dtype, small model, no GPU.

* large-separation additivity and rigid-translation invariance          model_tests_basic.py:507-594
* E(3) + permutation equivariance of energies and forces                :450-461 (assert_AtomicData_equivariant)
* pair force: non-zero inside, exactly zero at and beyond the cutoff    :811-843
* edge embedding vanishes at the cutoff, other edges unaffected         :964-1000
* isolated atoms restore the per-type energy shifts                     :932-962
"""
import math

import numpy as np
import pytest
import torch

from nequip_b200 import data as D
from nequip_b200.nn.model import NequIPEnergyModel
from oracle import model as omodel

R_MAX = 4.0
TYPES = ["H", "C", "O"]


@pytest.fixture(scope="module")
def net():
    m = NequIPEnergyModel(r_max=R_MAX, type_names=TYPES, l_max=2, num_layers=3, num_features=8, radial_mlp_width=16,
                          avg_num_neighbors=9.0, per_type_energy_scales=[1.3, 0.7, 2.1],
                          per_type_energy_shifts=[-0.5, 3.25, 11.0], model_dtype=torch.float64, seed=7)
    return m.state_dict(), m.config


def _cluster(n, seed, spread=2.2):
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, 3)) * spread
    types = rng.integers(0, len(TYPES), size=n)
    return pos, types


def _frame(pos, types):
    ei, _ = D.neighbor_list(np.asarray(pos, dtype=np.float64), None, R_MAX)
    return {"pos": torch.as_tensor(pos, dtype=torch.float64), "atom_types": torch.as_tensor(types, dtype=torch.long),
            "edge_index": torch.from_numpy(ei)}


def _run(net, frame):
    sd, cfg = net
    return omodel.energy_and_forces(sd, cfg, frame, torch.float64)


def test_large_separation_additivity_and_rigid_translation(net):
    p1, t1 = _cluster(7, 1)
    p2, t2 = _cluster(6, 2)
    rng = np.random.default_rng(3)
    p2 = p2 + 40.0 + rng.normal(size=3)
    f1, f2 = _frame(p1, t1), _frame(p2, t2)
    both = _frame(np.concatenate([p1, p2]), np.concatenate([t1, t2]))
    assert both["edge_index"].shape[1] == f1["edge_index"].shape[1] + f2["edge_index"].shape[1] > 0
    e1, a1, g1 = _run(net, f1)
    e2, a2, g2 = _run(net, f2)
    eb, ab, gb = _run(net, both)
    assert torch.allclose(e1 + e2, eb, atol=1e-10)
    assert torch.allclose(torch.cat([g1, g2]), gb, atol=1e-10)
    # rigid translation of the second molecule: total and per-atom energies unchanged
    both2 = _frame(np.concatenate([p1, p2 + rng.normal(size=3)]), np.concatenate([t1, t2]))
    eb2, ab2, _ = _run(net, both2)
    assert torch.allclose(eb2, eb, atol=1e-10) and torch.allclose(ab2, ab, atol=1e-10)


@pytest.mark.parametrize("improper", [False, True])
def test_energy_and_forces_are_E3_and_permutation_equivariant(net, improper):
    pos, types = _cluster(12, 5, spread=1.8)
    e0, a0, f0 = _run(net, _frame(pos, types))
    g = torch.Generator().manual_seed(11)
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=g))
    R = q * torch.sign(torch.linalg.det(q))  # a proper rotation
    if improper:
        R = -R
    perm = torch.randperm(pos.shape[0], generator=g)
    shift = torch.randn(3, dtype=torch.float64, generator=g) * 5
    pos2 = (torch.as_tensor(pos) @ R.t() + shift)[perm]
    e1, a1, f1 = _run(net, _frame(pos2.numpy(), np.asarray(types)[perm.numpy()]))
    assert torch.allclose(e1, e0, atol=1e-10)
    assert torch.allclose(a1, a0[perm], atol=1e-10)
    assert torch.allclose(f1, (f0 @ R.t())[perm], atol=1e-10)  # forces are polar vectors: F -> R F, also for det R = -1


def _pair_forces(net, ti, tj, dist, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.normal(size=3)
    u /= np.linalg.norm(u)
    frame = {"pos": torch.as_tensor(np.stack([np.zeros(3), dist * u])), "atom_types": torch.tensor([ti, tj]),
             "edge_index": torch.tensor([[0, 1], [1, 0]])}  # the edge exists whatever the distance (list cutoff 1.5 r_max)
    return _run(net, frame)[2]


def test_pair_force_is_zero_at_and_beyond_the_cutoff(net):
    for ti in range(len(TYPES)):
        for tj in range(len(TYPES)):
            assert _pair_forces(net, ti, tj, 0.5 * R_MAX).abs().sum() > 1e-4  # control: interacting inside
            for d in (R_MAX, 1.1 * R_MAX):
                f = _pair_forces(net, ti, tj, d)
                assert torch.allclose(f, torch.zeros_like(f)), (ti, tj, d, f)


def test_edge_embedding_vanishes_at_the_cutoff():
    pos = torch.tensor([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 1, 0, 2], [1, 0, 2, 0]])
    _, _, emb = omodel.edge_embed(pos, ei, None, None, 2, 8, R_MAX, 6.0, torch.float64)
    pos2 = pos.clone()
    pos2[2, 1] = R_MAX  # put it at the cutoff
    _, _, emb2 = omodel.edge_embed(pos2, ei, None, None, 2, 8, R_MAX, 6.0, torch.float64)
    torch.testing.assert_close(emb[:2], emb2[:2])          # other edges unaffected
    assert emb[2:].abs().sum() > 1e-6                      # non-zero before
    torch.testing.assert_close(emb2[2:], torch.zeros_like(emb2[2:]))
    # and its derivative with respect to the positions vanishes there too (smooth envelope, p = 6)
    p = pos2.clone().requires_grad_(True)
    _, _, e3 = omodel.edge_embed(p, ei, None, None, 2, 8, R_MAX, 6.0, torch.float64)
    (g,) = torch.autograd.grad(e3[2:].sum(), p)
    assert torch.isfinite(g).all() and g.abs().max() < 1e-12


def test_isolated_atoms_restore_the_per_type_shifts(net):
    sd, cfg = net
    for t in range(len(TYPES)):
        frame = {"pos": torch.zeros(1, 3, dtype=torch.float64), "atom_types": torch.tensor([t]),
                 "edge_index": torch.zeros(2, 0, dtype=torch.long)}
        e, a, f = omodel.energy_and_forces(sd, cfg, frame, torch.float64)
        # no neighbours -> no message: conv output is the self-connection only, which the readout maps to ...
        # the reference's statement (model_tests_basic.py:932-962) holds for models whose untrained readout is zero
        # (zero_prior_at_init); in general E = scale * readout(features of an isolated atom) + shift:
        assert torch.isfinite(e).all() and f.abs().max() == 0
        e_scale_free = (float(e) - float(sd["shifts"][t])) / float(sd["scales"][t])
        # the same atom with every scale = 1 and shift = 0 gives exactly that residual
        sd0 = dict(sd)
        sd0["scales"], sd0["shifts"] = torch.ones_like(sd["scales"]), torch.zeros_like(sd["shifts"])
        e_plain, _, _ = omodel.energy_and_forces(sd0, cfg, frame, torch.float64)
        assert abs(float(e_plain) - e_scale_free) < 1e-12
    # with a zero readout (the reference's zero-prior initialisation) the shifts are restored exactly
    sdz = dict(sd)
    sdz["readout.mlp.0.weight"] = torch.zeros_like(sd["readout.mlp.0.weight"])
    for t in range(len(TYPES)):
        frame = {"pos": torch.zeros(1, 3, dtype=torch.float64), "atom_types": torch.tensor([t]),
                 "edge_index": torch.zeros(2, 0, dtype=torch.long)}
        e, _, _ = omodel.energy_and_forces(sdz, cfg, frame, torch.float64)
        assert float(e) == pytest.approx(float(sd["shifts"][t]), abs=1e-14)


# ---- geometry: tests/unit/nn/test_utils.py:15-89 and model_tests_basic.py:326-383 (wrapped / unwrapped images)
def _fcc(a=3.61, reps=1):
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    cells = np.array([(i, j, k) for i in range(reps) for j in range(reps) for k in range(reps)]) * a
    return (base[None] + cells[:, None]).reshape(-1, 3), np.eye(3) * a * reps


def test_periodic_edges_of_close_packed_bulk():
    """test_utils.py:32-42: every atom of fcc bulk has 12 neighbours at the nearest-neighbour distance."""
    pos, cell = _fcc()
    dist = 3.61 / math.sqrt(2)
    ei, sh = D.neighbor_list(pos, cell, 1.05 * dist)
    vec = omodel.edge_vectors(torch.from_numpy(pos), torch.from_numpy(ei), torch.from_numpy(cell), torch.from_numpy(sh))
    assert ei.shape[1] == 12 * pos.shape[0] and (np.bincount(ei[0]) == 12).all()
    torch.testing.assert_close(vec.norm(dim=-1), torch.full((ei.shape[1],), dist, dtype=torch.float64))
    # gradients with respect to positions and cell exist (test_utils.py:45-69)
    p, c = torch.from_numpy(pos).requires_grad_(True), torch.from_numpy(cell).requires_grad_(True)
    v = omodel.edge_vectors(p, torch.from_numpy(ei), c, torch.from_numpy(sh))
    gp, gc = torch.autograd.grad(v.square().sum(), [p, c])  # (the plain sum cancels between the two directions of an edge)
    assert torch.isfinite(gp).all() and torch.isfinite(gc).all() and gc.abs().sum() > 0


def test_wrapped_and_unwrapped_periodic_images_give_the_same_result(net):
    """model_tests_basic.py:326-383: moving atoms into other periodic images changes the edge shifts by
    ``cs[centre] - cs[neighbour]`` and nothing else."""
    sd, cfg = net
    pos, cell = _fcc(3.9, reps=2)  # 32 atoms, box 7.8 A < 2 r_max: several images per pair
    rng = np.random.default_rng(12345)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    types = rng.integers(0, len(TYPES), size=pos.shape[0])
    ei, sh = D.neighbor_list(pos, cell, 3.5)
    frame = {"pos": torch.from_numpy(pos), "cell": torch.from_numpy(cell), "atom_types": torch.from_numpy(types),
             "edge_index": torch.from_numpy(ei), "edge_cell_shift": torch.from_numpy(sh)}
    e0, a0, f0 = omodel.energy_and_forces(sd, cfg, frame, torch.float64)
    for _ in range(3):
        cs = rng.integers(-5, 5, size=(pos.shape[0], 3)).astype(np.float64)
        pos2 = pos + cs @ cell
        ei2, sh2 = D.neighbor_list(pos2, cell, 3.5)
        assert np.array_equal(ei, ei2)
        assert np.array_equal(sh + cs[ei[0]] - cs[ei[1]], sh2)
        frame2 = dict(frame, pos=torch.from_numpy(pos2), edge_cell_shift=torch.from_numpy(sh2))
        e1, a1, f1 = omodel.energy_and_forces(sd, cfg, frame2, torch.float64)
        assert torch.allclose(e1, e0, atol=1e-9) and torch.allclose(a1, a0, atol=1e-9) and torch.allclose(f1, f0, atol=1e-9)


def test_host_neighbor_list_handles_unwrapped_positions_cell_list_path():
    """Same contract on the cell-list path of the host neighbour list (>= 64 atoms, box >= 3 r_max)."""
    pos, cell = D.jittered_lattice(8, D.PRESETS["li3po4"]["density"], seed=3)
    ei, sh = D.neighbor_list(pos, cell, 5.0)
    cs = np.random.default_rng(0).integers(-5, 5, size=(pos.shape[0], 3)).astype(np.float64)
    ei2, sh2 = D.neighbor_list(pos + cs @ cell, cell, 5.0)
    assert np.array_equal(ei, ei2) and np.array_equal(sh + cs[ei[0]] - cs[ei[1]], sh2)


def test_batched_frames_equal_the_frames_evaluated_one_by_one(net):
    """model_tests_basic.py:385-448 (``test_batch``) and :598-629 (cross-frame gradient isolation): graph fields per
    frame, node fields per atom, and no force leaks from one frame of a batch into another."""
    sd, cfg = net
    frames = [_frame(*_cluster(n, seed)) for n, seed in ((6, 21), (9, 22), (1, 23))]
    outs = [omodel.energy_and_forces(sd, cfg, f, torch.float64) for f in frames]
    off = np.cumsum([0] + [f["pos"].shape[0] for f in frames])
    batched = {
        "pos": torch.cat([f["pos"] for f in frames]),
        "atom_types": torch.cat([f["atom_types"] for f in frames]),
        "edge_index": torch.cat([f["edge_index"] + int(o) for f, o in zip(frames, off)], dim=1),
        "batch": torch.cat([torch.full((f["pos"].shape[0],), i) for i, f in enumerate(frames)]),
        "num_atoms": torch.tensor([f["pos"].shape[0] for f in frames]),
    }
    e, a, f = omodel.energy_and_forces(sd, cfg, batched, torch.float64)
    assert e.shape == (3, 1)
    for i, (ei, ai, fi) in enumerate(outs):
        sl = slice(int(off[i]), int(off[i + 1]))
        assert torch.allclose(e[i], ei.view(-1), atol=1e-10)
        assert torch.allclose(a[sl], ai, atol=1e-10) and torch.allclose(f[sl], fi, atol=1e-10)
    # cross-frame isolation: the energy of frame 0 has no gradient on the atoms of the other frames
    pos = batched["pos"].clone().requires_grad_(True)
    e_tot, _ = omodel.energy(sd, cfg, dict(batched, pos=pos), torch.float64)
    (g0,) = torch.autograd.grad(e_tot[0].sum(), pos)
    assert g0[: int(off[1])].abs().sum() > 0 and float(g0[int(off[1]):].abs().max()) == 0.0


def test_mixed_boundary_conditions_monolayer():
    """test_utils.py:72-95 (``test_some_periodic``): an fcc(111) monolayer periodic in x and y only -- every atom has
    its 6 in-plane first-shell neighbours and no edge has a z component (no images across the vacuum)."""
    a_nn = 2.864  # Al first-shell distance
    # orthorhombic 2-atom cell of the triangular lattice, 3 x 2 cells, 20 A of vacuum along z
    base = np.array([[0.0, 0.0, 10.0], [0.5 * a_nn, 0.5 * math.sqrt(3) * a_nn, 10.0]])
    pos = np.concatenate([base + np.array([i * a_nn, j * math.sqrt(3) * a_nn, 0.0]) for i in range(3) for j in range(2)])
    cell = np.diag([3 * a_nn, 2 * math.sqrt(3) * a_nn, 20.0])
    ei, sh = D.neighbor_list(pos, cell, 2.9, pbc=(True, True, False))
    assert (np.bincount(ei[0], minlength=pos.shape[0]) == 6).all()
    vec = omodel.edge_vectors(torch.from_numpy(pos), torch.from_numpy(ei), torch.from_numpy(cell), torch.from_numpy(sh))
    torch.testing.assert_close(vec[:, 2], torch.zeros(ei.shape[1], dtype=torch.float64))
    torch.testing.assert_close(vec.norm(dim=-1), torch.full((ei.shape[1],), a_nn, dtype=torch.float64))
    assert (sh[:, 2] == 0).all()
    # fully periodic with a z length below r_max would add images across z: the flags matter
    thin = np.diag([3 * a_nn, 2 * math.sqrt(3) * a_nn, 2.5])
    ei_z, _ = D.neighbor_list(pos, thin, 2.9, pbc=True)
    ei_slab, _ = D.neighbor_list(pos, thin, 2.9, pbc=(True, True, False))
    assert ei_z.shape[1] > ei_slab.shape[1] == ei.shape[1]
