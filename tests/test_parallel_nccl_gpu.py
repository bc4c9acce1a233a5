"""2 GPUs, NCCL: one frame split into two slabs with a per-layer halo exchange
(nequip_b200/parallel.py; reference design nequip/nn/_ghost_exchange_base.py:8-57,
nequip/nn/interaction_block.py:159-199) must give the energy and forces of the unsharded model on the
same frame (fp32 kernels; 1e-5 relative).  Skipped on a box with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nequip_b200 import data as D
from nequip_b200 import parallel as P
from nequip_b200.nn.model import NequIPEnergyModel

pytestmark = pytest.mark.gpu

MK = dict(l_max=2, num_layers=4, num_features=64, radial_mlp_depth=1, radial_mlp_width=128)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(meta, dev):
    torch.manual_seed(123)
    m = NequIPEnergyModel(r_max=5.0, type_names=meta["type_names"], parity=True,
                          avg_num_neighbors=meta["avg_num_neighbors"], **MK).to(dev)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def _worker(rank, world, port, sysd, meta, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model = _model(meta, dev)
        owner = P.slab_owner(sysd["pos"], world)
        plan = P.make_plans(sysd["edge_index"], owner, world)[rank]
        local = D.to_device(P.shard_data(sysd, plan), dev)
        halo = P.HaloExchange(plan, dev)
        e, f = P.sharded_energy_forces(model, local, plan, halo)
        torch.cuda.synchronize()
        if rank == 0:
            ret["e"], ret["f"] = e.cpu(), f.cpu()
            ret["ghosts"] = plan.n_ghost
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_halo_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sysd = D.make_system("li3po4", 8, r_max=5.0, seed=2)
    meta = sysd.pop("_meta")
    ref = _model(meta, "cuda:0")(D.to_device(sysd, "cuda:0"))
    e_ref, f_ref = ref["total_energy"].cpu(), ref["forces"].cpu()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), sysd, meta, ret), nprocs=2, join=True)
    assert ret["ghosts"] > 0
    escale = float(ref["atomic_energy"].abs().sum())
    assert abs(float(ret["e"]) - float(e_ref)) <= 1e-5 * escale
    assert float((ret["f"] - f_ref).abs().max()) <= 5e-5 * float(f_ref.abs().max())
