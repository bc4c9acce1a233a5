"""GPU parity of the fused "last radial-MLP layer -> tensor product -> scatter" forward kernel
(nqb_tp_fused_fwd, nequip_b200/csrc/nqb_tp_fused.cuh) against the unfused composition
  w = h @ (W2 * alpha2)              nequip/nn/mlp.py:262-268
  out = scatter(tp(x[src], Y, w))    nequip/nn/_tp_scatter_base.py:35-38
(the unfused kernels are themselves checked against the oracle in test_tp_scatter_gpu.py), and of the whole
model with and without the fused kernel against the oracle."""
import math

import pytest
import torch

from nequip_b200 import data as D
from nequip_b200 import known_signatures as ks
from nequip_b200 import ops
from nequip_b200.codegen import GenOptions, TPGenerator
from nequip_b200.nn.model import NequIPEnergyModel

pytestmark = pytest.mark.gpu


def _graph(N, degs, seed):
    """Edges grouped by destination with the given per-node degrees (0, 1, odd, > 64 ...)."""
    g = torch.Generator().manual_seed(seed)
    dst = torch.repeat_interleave(torch.arange(N), torch.tensor(degs))
    src = torch.randint(0, N, (dst.numel(),), generator=g)
    return dst, src


CASES = [
    # (l_max, features, layer index, node degrees)
    (2, 64, 2, [0, 1, 2, 3, 64, 65, 0, 0, 150, 17, 16, 15, 33, 7]),
    (2, 64, 1, [5, 0, 130, 64, 63]),
    (2, 64, 0, [3, 70, 1]),
    (2, 64, 3, [9, 0, 66]),
    (2, 32, 2, [0, 1, 2, 3, 64, 65, 0, 129, 31]),
    (2, 32, 1, [12, 77, 1]),
    (3, 32, 2, [4, 0, 66, 13]),
    (3, 32, 3, [1, 65, 20]),
]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("lmax,nf,li,degs", CASES)
@pytest.mark.parametrize("K", [128, 64])
def test_fused_forward_matches_unfused(lmax, nf, li, degs, K):
    nl = 5 if lmax == 3 else 4
    sig = ks.nequip_layer_signatures(lmax, nf, nl)[li]
    opts = GenOptions(layout="ir_mul")
    if TPGenerator(sig, opts).fused_layout() is None:
        pytest.skip("signature has no fused kernel")
    plan = ops.get_plan(sig.irreps_in1, sig.irreps_in2, sig.irreps_out, sig.instructions, opts)
    N = len(degs)
    dst, src = _graph(N, degs, seed=li + lmax)
    E = dst.numel()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, sig.d_in, generator=g).cuda()
    y = torch.randn(E, sig.s_dim, generator=g).cuda()
    h = torch.randn(E, K, generator=g).cuda()
    W2 = ((torch.rand(K, sig.weight_numel, generator=g) * 2 - 1) * math.sqrt(3)).cuda()
    a2 = math.sqrt(2) / math.sqrt(K)
    fw = ops.FusedTPWeights(plan, W2, a2, "cuda")
    csr = ops.build_csr(dst.cuda(), N)
    out, w = ops.tp_fused_fwd(fw, x, y, h, src.cuda(), csr, want_w=True)
    out2, w2 = ops.tp_fused_fwd(fw, x, y, h, src.cuda(), csr, want_w=False)
    torch.cuda.synchronize()
    assert w2 is None and torch.equal(out, out2)  # deterministic, independent of the side output
    w_ref = (h.double() @ (W2.double() * a2))
    werr = (w.double() - w_ref).abs().max().item() / w_ref.abs().max().item()
    assert werr <= 2e-6, werr
    out_ref = ops.tp_scatter(plan, x.double(), y.double(), w_ref, dst.cuda(), src.cuda(), csr=csr)  # fp64 kernels
    err = (out.double() - out_ref).abs().max().item() / out_ref.abs().max().item()
    assert err <= 3e-6, err
    # isolated nodes are written as zeros
    iso = torch.tensor([d == 0 for d in degs])
    assert float(out[iso.cuda()].abs().max() if iso.any() else 0.0) == 0.0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name,kind,mk", [
    ("li3po4_l2_f64feat", "li3po4", dict(l_max=2, num_layers=4, num_features=64, radial_mlp_depth=1, radial_mlp_width=128)),
    ("water_l2_f32", "water", dict(l_max=2, num_layers=4, num_features=32, radial_mlp_depth=1, radial_mlp_width=128)),
    ("asi_l3", "asi", dict(l_max=3, num_layers=5, num_features=32, radial_mlp_depth=1, radial_mlp_width=128)),
    ("water_l2_w64", "water", dict(l_max=2, num_layers=3, num_features=32, radial_mlp_depth=1, radial_mlp_width=64)),
])
def test_model_with_fused_kernel_matches_oracle_and_unfused(name, kind, mk):
    from oracle import model as omodel

    sysd = D.make_system(kind, 6, r_max=5.0, seed=1)
    meta = sysd.pop("_meta")
    model = NequIPEnergyModel(r_max=5.0, type_names=meta["type_names"], parity=True,
                              avg_num_neighbors=meta["avg_num_neighbors"], strict_fast_path=True, **mk).cuda()
    for p in model.parameters():
        p.requires_grad_(False)
    for l in model.layers:
        l.conv.use_fused_radial_tp = True  # (the default "auto" times both paths once and keeps the faster)
    dev = D.to_device(sysd, "cuda")
    out = model(dev)
    used = [l.conv._tc_cache[1]["fused"] is not None for l in model.layers]
    assert any(used), "no layer used the fused kernel"
    for l in model.layers:
        l.conv.use_fused_radial_tp = False
    ref = model(dev)
    fscale = float(ref["forces"].abs().max())
    assert float((out["forces"] - ref["forces"]).abs().max()) <= 2e-6 * fscale
    assert abs(float(out["total_energy"]) - float(ref["total_energy"])) <= 2e-6 * float(ref["atomic_energy"].abs().sum())
    e_ref, ea_ref, f_ref = omodel.energy_and_forces(model.state_dict(), model.config, sysd, torch.float32)
    e, f = out["total_energy"].cpu(), out["forces"].cpu()
    assert abs(float(e) - float(e_ref)) <= 1e-5 * float(ea_ref.abs().sum())
    assert float((f - f_ref).abs().max()) <= 1e-5 * float(f_ref.abs().max())
