"""CPU, world_size 2, gloo: the spatial-decomposition plumbing of nequip_b200.parallel
(owned/ghost numbering, halo plans, forward exchange and its transposed backward, energy/force
reductions) on a toy message-passing energy written in plain torch.  The sharded result must equal
the single-process result including gradients -- the property the reference's LAMMPS ghost
exchange (nequip/nn/_ghost_exchange_lmp_mliap.py:11-64) has to satisfy."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nequip_b200 import data as D
from nequip_b200 import parallel as P


def _toy_energy(pos, types, edge_index, cell, shift, n_own, halo, layers=3):
    """x_{l+1}[i] = tanh( sum_{j in N(i)} f(|r_ij|) * x_l[j] ) (+ ghost refresh per layer), E = sum_i x_L[i]."""
    vec = pos[edge_index[1]] - pos[edge_index[0]] + shift @ cell
    r = vec.norm(dim=1, keepdim=True)
    f = torch.cos(r) / (1 + r)
    x = torch.stack([torch.sin(types.double() + 1.0), torch.cos(types.double() * 0.5)], 1)  # "embedding" (all local atoms)
    for l in range(layers):
        if l > 0:
            x = halo(x[:n_own])
        msg = torch.zeros_like(x).index_add(0, edge_index[0], f * x[edge_index[1]])
        x = torch.tanh(msg + 0.1 * x)[:n_own]
    return x.sum(dim=1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _retry_once(test):
    """Multi-process CPU tests share the machine with whatever else runs on it (the suite has been seen to lose ~1 run
    in 30 to a transient rendezvous / scheduling hiccup): a failed test is run once more; a genuine failure fails twice."""
    import functools

    @functools.wraps(test)
    def wrapped(*a, **k):
        try:
            return test(*a, **k)
        except Exception:  # noqa: BLE001
            return test(*a, **k)

    return wrapped


def _spawn(worker, world, args_after_port):
    """``mp.spawn(worker, (world, port, *args))`` with one retry on a fresh port: the port is probed before the
    workers bind it, and a rendezvous that loses that race (or a transient gloo connect error) must not fail the suite.
    A genuine failure fails again."""
    last = None
    for _ in range(2):
        try:
            mp.spawn(worker, args=(world, _free_port(), *args_after_port), nprocs=world, join=True)
            return
        except Exception as exc:  # noqa: BLE001 - re-raised after the retry
            last = exc
    raise last


def _worker(rank, world, port, sysd, ret, grid=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = P.slab_owner(sysd["pos"], world) if grid is None else P.brick_owner(sysd["pos"], grid)
        plan = P.make_plans(sysd["edge_index"], owner, world)[rank]
        local = P.shard_data(sysd, plan)
        halo = P.HaloExchange(plan, "cpu")
        pos = local["pos"].clone().requires_grad_(True)
        e_own = _toy_energy(pos, local["atom_types"], local["edge_index"], local["cell"], local["edge_cell_shift"],
                            plan.n_own, halo)
        e_loc = e_own.sum()
        (g,) = torch.autograd.grad(e_loc, pos)
        e = e_loc.detach().reshape(1).clone()
        dist.all_reduce(e)
        f = torch.zeros(plan.num_global, 3, dtype=torch.float64)
        f.index_add_(0, plan.local_ids, -g)
        dist.all_reduce(f)
        # owner reduction (what the product uses at scale): ghost gradients travel back to their owners
        f_own = -P.owner_reduce(g, plan, halo)
        assert f_own.shape == (plan.n_own, 3)
        torch.testing.assert_close(f_own, f[plan.owned], atol=1e-12, rtol=1e-10)
        if rank == 0:
            ret["e"], ret["f"] = e, f
            ret["ghost_frac"] = plan.n_ghost / max(plan.n_own, 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,grid", [(2, None), (3, None), (4, (2, 2, 1))])
@_retry_once
def test_sharded_toy_model_matches_single_process(world, grid):
    sysd = D.make_system("water", 6, r_max=5.0, seed=4)
    sysd.pop("_meta")
    # single process reference (world 1: halo is the identity)
    plan1 = P.make_plans(sysd["edge_index"], torch.zeros(sysd["pos"].shape[0], dtype=torch.long), 1)[0]
    halo1 = P.HaloExchange(plan1, "cpu")
    pos = sysd["pos"].clone().requires_grad_(True)
    e_ref = _toy_energy(pos, sysd["atom_types"], sysd["edge_index"], sysd["cell"], sysd["edge_cell_shift"],
                        pos.shape[0], halo1).sum()
    (g_ref,) = torch.autograd.grad(e_ref, pos)
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_worker, world, (sysd, ret, grid,))
    assert abs(float(ret["e"]) - float(e_ref)) < 1e-10 * max(1.0, abs(float(e_ref)))
    torch.testing.assert_close(ret["f"], -g_ref.detach(), atol=1e-11, rtol=1e-9)
    assert ret["ghost_frac"] > 0


def test_plans_are_consistent():
    sysd = D.make_system("li3po4", 6, r_max=5.0, seed=1)
    world = 4
    owner = P.slab_owner(sysd["pos"], world)
    plans = P.make_plans(sysd["edge_index"], owner, world)
    N, E = sysd["pos"].shape[0], sysd["edge_index"].shape[1]
    assert sum(p.n_own for p in plans) == N and sum(p.edge_index.shape[1] for p in plans) == E
    for r, p in enumerate(plans):
        assert int(p.edge_index[0].max()) < p.n_own  # every destination is owned
        assert torch.equal(owner[p.owned], torch.full((p.n_own,), r))
        assert not (owner[p.ghosts] == r).any()
        for s, q in enumerate(plans):  # what r receives from s is what s sends to r
            assert p.recv_splits[s] == q.send_splits[r]
        # edges are the global edges with owned destination, re-indexed
        ge = sysd["edge_index"][:, p.edge_ids]
        assert torch.equal(p.local_ids[p.edge_index[0]], ge[0]) and torch.equal(p.local_ids[p.edge_index[1]], ge[1])


def test_brick_decomposition():
    sysd = D.make_system("li3po4", 8, r_max=5.0, seed=1)
    pos = sysd["pos"]
    assert torch.equal(P.brick_owner(pos, (4, 1, 1)), P.slab_owner(pos, 4))
    owner = P.brick_owner(pos, (2, 2, 2))
    assert torch.bincount(owner, minlength=8).tolist() == [64] * 8
    plans = P.make_plans(sysd["edge_index"], owner, 8)
    assert sum(p.n_own for p in plans) == pos.shape[0]
    # elongated boxes are cut into slabs, cubic boxes into bricks (fewest ghosts for a 5 A halo)
    assert P.brick_grid(8, [374.0, 46.8, 46.8]) == (8, 1, 1)
    g = P.brick_grid(8, [93.6, 93.6, 93.6])
    assert g[0] * g[1] * g[2] == 8 and max(g) < 8


def _worker_supercell(rank, world, port, full, n_base, f_base, e_base, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lengths = torch.diagonal(full["cell"]).tolist()
        grid = P.brick_grid(world, lengths, halo=5.0)
        owner = P.brick_owner(full["pos"], grid)
        plan = P.make_plans(full["edge_index"], owner, world)[rank]
        local = P.shard_data(full, plan)
        halo = P.HaloExchange(plan, "cpu")
        pos = local["pos"].clone().requires_grad_(True)
        e_loc = _toy_energy(pos, local["atom_types"], local["edge_index"], local["cell"], local["edge_cell_shift"],
                            plan.n_own, halo).sum()
        (g,) = torch.autograd.grad(e_loc, pos)
        e = e_loc.detach().reshape(1).clone()
        dist.all_reduce(e)
        f_own = -P.owner_reduce(g, plan, halo)
        # bench.py's partition-parity check: owned atom g is a periodic copy of base atom g % n_base
        err = (f_own - f_base[plan.owned % n_base]).abs().max().reshape(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            ret["err"], ret["e"], ret["grid"] = float(err), float(e), grid
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
@_retry_once
def test_supercell_partition_reproduces_the_base_frame(world):
    """bench.py's weak-scaling frame is the ``world``-fold periodic supercell of the N = 1 frame
    (``data.replicate_frame``): sharded over ``world`` ranks it must give every copy of an atom the force of the
    base atom in the unsharded base frame, and ``world`` times its energy."""
    base = D.make_system("water", 6, r_max=5.0, seed=4)
    base.pop("_meta")
    n = base["pos"].shape[0]
    full = D.replicate_frame(base, world, r_max=5.0)
    assert full["pos"].shape[0] == world * n and full["edge_index"].shape[1] == world * base["edge_index"].shape[1]
    assert torch.equal(full["atom_types"][n: 2 * n], base["atom_types"])
    plan1 = P.make_plans(base["edge_index"], torch.zeros(n, dtype=torch.long), 1)[0]
    pos = base["pos"].clone().requires_grad_(True)
    e_base = _toy_energy(pos, base["atom_types"], base["edge_index"], base["cell"], base["edge_cell_shift"], n,
                         P.HaloExchange(plan1, "cpu")).sum()
    (g_base,) = torch.autograd.grad(e_base, pos)
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_worker_supercell, world, (full, n, -g_base.detach(), float(e_base), ret,))
    assert ret["grid"] == (world, 1, 1)  # elongated along x -> slabs
    assert ret["err"] < 1e-10 * float(g_base.abs().max())
    assert abs(ret["e"] - world * float(e_base)) < 1e-10 * abs(world * float(e_base))


def _load_bench():
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("nqb_bench_checks", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _toy_unsharded(frame):
    """An 'unsharded model call' with the output keys of NequIPEnergyModel.forward, on the toy energy."""
    n = frame["pos"].shape[0]
    plan1 = P.make_plans(frame["edge_index"], torch.zeros(n, dtype=torch.long), 1)[0]
    pos = frame["pos"].clone().requires_grad_(True)
    e_atom = _toy_energy(pos, frame["atom_types"], frame["edge_index"], frame["cell"], frame["edge_cell_shift"], n,
                         P.HaloExchange(plan1, "cpu"))
    (g,) = torch.autograd.grad(e_atom.sum(), pos)
    return {"forces": -g, "total_energy": e_atom.sum().detach().reshape(1, 1), "atomic_energy": e_atom.detach().reshape(-1, 1)}


def _worker_checks(rank, world, port, full, base, copies, break_it, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bench = _load_bench()
        grid = P.brick_grid(world, torch.diagonal(full["cell"]).tolist(), halo=5.0)
        plan = P.make_plans(full["edge_index"], P.brick_owner(full["pos"], grid), world)[rank]
        local = P.shard_data(full, plan)
        halo = P.HaloExchange(plan, "cpu")

        def step():  # what bench.py times in halo mode: sharded energy + owner-reduced forces
            pos = local["pos"].clone().requires_grad_(True)
            e_loc = _toy_energy(pos, local["atom_types"], local["edge_index"], local["cell"], local["edge_cell_shift"],
                                plan.n_own, halo).sum()
            (g,) = torch.autograd.grad(e_loc, pos)
            e = e_loc.detach().reshape(1).clone()
            dist.all_reduce(e)
            f = -P.owner_reduce(g, plan, halo)
            if break_it == 1 and rank == 1:
                f = -g[: plan.n_own]  # a rank that forgets the ghost contributions it owes to / is owed by others
            return {"total_energy": e, "forces": f}

        model = _toy_unsharded
        if break_it == 2 and rank == 0:
            def model(frame):  # a rank whose local check blows up: reported, the collectives still line up
                raise RuntimeError("boom")
        checks = bench.parity_checks(step, model, base, copies, plan.owned, True, world, rank, torch.device("cpu"))
        if rank == 0:
            ret["checks"] = checks
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("break_it", [0, 1, 2])
@_retry_once
def test_bench_parity_checks_under_gloo(break_it):
    """bench.py's ``parity_checks`` executed for real on 2 ranks (toy energy): green for a correct sharded step, a
    dropped ghost contribution is flagged by both properties, and a rank whose local check raises is reported
    without desynchronising the collectives."""
    world = 2
    base = D.make_system("water", 6, r_max=5.0, seed=4)
    base.pop("_meta")
    full = D.replicate_frame(base, world, r_max=5.0)
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_worker_checks, world, (full, base, world, break_it, ret,))
    c = ret["checks"]
    pp = c["partition_parity"]
    assert c["ranks_reporting"] == world
    if break_it == 0:
        assert c["sum_forces_over_sum_abs_forces"] < 1e-12
        assert pp["ranks_reporting"] == world and pp["max_dF_over_max_F"] < 1e-10 and pp["dE_over_sum_abs_Ei"] < 1e-12
    elif break_it == 1:
        assert c["sum_forces_over_sum_abs_forces"] > 1e-4 and pp["max_dF_over_max_F"] > 1e-3
    else:
        assert pp["ranks_reporting"] == world - 1 and pp["max_dF_over_max_F"] < 1e-10


def test_bench_parity_checks_single_rank():
    bench = _load_bench()
    base = D.make_system("water", 5, r_max=5.0, seed=1)
    base.pop("_meta")
    c = bench.parity_checks(lambda: _toy_unsharded(base), None, None, 1, None, False, 1, 0, torch.device("cpu"))
    assert "partition_parity" not in c and c["ranks_reporting"] == 1 and c["sum_forces_over_sum_abs_forces"] < 1e-12


def _worker_exchange_profile(rank, world, port, full, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bench = _load_bench()
        grid = P.brick_grid(world, torch.diagonal(full["cell"]).tolist(), halo=5.0)
        plan = P.make_plans(full["edge_index"], P.brick_owner(full["pos"], grid), world)[rank]
        prof = bench.halo_exchange_profile([(1, 12), (2, 40)], plan, P.HaloExchange(plan, "cpu"), "cpu", world, reps=2)
        if rank == 0:
            ret["prof"], ret["sent"], ret["ghost"] = prof, int(sum(plan.send_splits)), plan.n_ghost
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@_retry_once
def test_bench_halo_exchange_profile_under_gloo():
    """bench.py's timing of the data-path collective (per-layer halo exchange, forward and forward + backward)."""
    base = D.make_system("water", 6, r_max=5.0, seed=4)
    base.pop("_meta")
    full = D.replicate_frame(base, 2, r_max=5.0)
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_worker_exchange_profile, 2, (full, ret,))
    prof = ret["prof"]
    assert [p["layer"] for p in prof] == [1, 2] and [p["row_floats"] for p in prof] == [12, 40]
    for p in prof:
        assert p["rows_sent"] == ret["sent"] and p["rows_received"] == ret["ghost"]
        assert p["bytes_sent_per_exchange"] == ret["sent"] * p["row_floats"] * 4
        assert p["ms_forward"] > 0 and p["ms_forward_plus_backward"] > 0
