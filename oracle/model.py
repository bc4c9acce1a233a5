"""CPU restatement of the NequIP energy model, e3nn formulation (TEST INFRASTRUCTURE ONLY).

Functional: ``energy_and_forces(state_dict, config, data)`` evaluates the same network
as ``nequip_b200.nn.model.NequIPEnergyModel`` from that module's ``state_dict`` -- in the
reference's own formulation (gather, per-path einsum with the w3j, scatter_add_,
autograd forces), in torch on the CPU.  It shares no code with the product.

Reference files followed (under /root/reference):
  with_edge_vectors_           nequip/nn/utils.py:68-118
  SphericalHarmonicEdgeAttrs   nequip/nn/embedding/_edge.py:193-198  (fp64 SH, cast to model dtype)
  EdgeLengthNormalizer         nequip/nn/embedding/_edge.py:65-80
  BesselEdgeLengthEncoding     nequip/nn/embedding/_edge.py:136-150
  PolynomialCutoff             nequip/nn/embedding/cutoffs.py:17-27
  ApplyFactor 2pi/r_max^2      nequip/nn/misc.py:46-48, nequip/model/nequip_models.py:318-322
  InteractionBlock.forward     nequip/nn/interaction_block.py:158-207
  ScalarMLPFunction            nequip/nn/mlp.py:133-195, 262-268
  AvgNumNeighborsNorm          nequip/nn/norm.py:48-68
  ConvNetLayer (gate)          nequip/nn/convnetlayer.py:74-170
  readout / scale-shift / sum  nequip/nn/mlp.py:75-77, nequip/nn/atomwise.py:236-284, :92-113
  ForceStressOutput            nequip/nn/grad_output.py:215-232
  e3nn o3.Linear / FullyConnectedTensorProduct / nn.Gate: SURVEY.md Appendix A.4 (e3nn 0.6.x, not vendored)
"""
import math
from typing import Dict

import torch

from . import irreps as I
from . import sh as osh
from . import tp as otp

C_SILU = 1.6791767923989418  # e3nn normalize2mom(silu)
C_TANH = 1.5937334472592692  # e3nn normalize2mom(tanh)


# ------------------------------------------------------------------ edge embedding
def edge_vectors(pos, edge_index, cell=None, shift=None):
    vec = torch.index_select(pos, 0, edge_index[1]) - torch.index_select(pos, 0, edge_index[0])
    if cell is not None:
        vec = vec + torch.sum(shift.view(-1, 3, 1) * cell.view(3, 3), 1)
    return vec


def polynomial_cutoff(x, p: float):
    out = 1.0
    out = out - (((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(x, p))
    out = out + (p * (p + 2.0) * torch.pow(x, p + 1.0))
    out = out - ((p * (p + 1.0) / 2) * torch.pow(x, p + 2.0))
    return out * (x < 1.0)


def radial_embedding(r, r_max: float, num_bessels: int, p: float, model_dtype):
    x = r.view(-1, 1) * (1.0 / r_max)
    bw = torch.linspace(1.0, num_bessels, num_bessels, dtype=torch.float64).unsqueeze(0)
    bessel = (torch.sinc(x * bw) * bw).to(model_dtype)
    cutoff = polynomial_cutoff(x, p).to(model_dtype)
    return ((2 * math.pi) / (r_max * r_max)) * (bessel * cutoff)


def edge_embed(pos, edge_index, cell, shift, lmax, num_bessels, r_max, p, model_dtype):
    vec = edge_vectors(pos, edge_index, cell, shift)
    r = vec.square().sum(1, keepdim=True).sqrt()
    y = osh.spherical_harmonics(lmax, vec, normalize=True).to(model_dtype)
    emb = radial_embedding(r, r_max, num_bessels, p, model_dtype)
    return vec, y, emb


# ------------------------------------------------------------------ dense e3nn pieces
def linear(x, weight, irreps_in, irreps_out):
    """o3.Linear: instructions (i_in, i_out) for equal irreps, i_in-major order;
    path weight 1/sqrt(sum of mul_in over paths into the same output)."""
    fin, fout = I.parse(irreps_in), I.parse(irreps_out)
    si = I.slices(fin)
    pairs = [(i, o) for i, (_, a) in enumerate(fin) for o, (_, b) in enumerate(fout) if a == b]
    N = x.shape[0]
    outs = [None] * len(fout)
    off = 0
    for (i, o) in pairs:
        mi, ir = fin[i]
        mo = fout[o][0]
        fan = sum(fin[i2][0] for (i2, o2) in pairs if o2 == o)
        W = weight[off: off + mi * mo].view(mi, mo)
        off += mi * mo
        xi = x[:, si[i]].reshape(N, mi, I.ir_dim(ir))
        r = torch.einsum("nuk,uw->nwk", xi, W) / math.sqrt(fan)
        r = r.reshape(N, -1)
        outs[o] = r if outs[o] is None else outs[o] + r
    for o, (mo, ir) in enumerate(fout):
        if outs[o] is None:
            outs[o] = x.new_zeros(N, mo * I.ir_dim(ir))
    return torch.cat(outs, 1)


def fctp_scalar_attr(x, attr, weight, irreps_in, num_attr, irreps_out):
    """FullyConnectedTensorProduct(irreps_in, num_attr x 0e, irreps_out), 'uvw'."""
    fin, fout = I.parse(irreps_in), I.parse(irreps_out)
    si = I.slices(fin)
    pairs = [(i, o) for i, (_, a) in enumerate(fin) for o, (_, b) in enumerate(fout) if a == b]
    N = x.shape[0]
    outs = [None] * len(fout)
    off = 0
    for (i, o) in pairs:
        mi, ir = fin[i]
        mo = fout[o][0]
        fan = sum(fin[i2][0] * num_attr for (i2, o2) in pairs if o2 == o)
        W = weight[off: off + mi * num_attr * mo].view(mi, num_attr, mo)
        off += mi * num_attr * mo
        xi = x[:, si[i]].reshape(N, mi, I.ir_dim(ir))
        r = torch.einsum("uvw,nuk,nv->nwk", W, xi, attr) / math.sqrt(fan)
        r = r.reshape(N, -1)
        outs[o] = r if outs[o] is None else outs[o] + r
    for o, (mo, ir) in enumerate(fout):
        if outs[o] is None:
            outs[o] = x.new_zeros(N, mo * I.ir_dim(ir))
    return torch.cat(outs, 1)


def _act(x, p):
    return torch.nn.functional.silu(x) * C_SILU if p == 1 else torch.tanh(x) * C_TANH


def gate(x, scalars, gates, gated):
    N = x.shape[0]
    parts, off = [], 0
    for mul, (l, p) in scalars:
        parts.append(_act(x[:, off: off + mul], p))
        off += mul
    gvals = []
    for mul, (l, p) in gates:
        gvals.append(_act(x[:, off: off + mul], p))
        off += mul
    if gated:
        g = torch.cat(gvals, 1)
        g0 = 0
        for mul, (l, p) in gated:
            d = 2 * l + 1
            ch = x[:, off: off + mul * d].reshape(N, mul, d)
            parts.append((ch * g[:, g0: g0 + mul].unsqueeze(-1)).reshape(N, mul * d))
            off += mul * d
            g0 += mul
    return torch.cat(parts, 1)


def tp_path_exists(in1, in2, ir_out):
    return any(ir_out in I.ir_mul(a, b) for _, a in I.simplify(in1) for _, b in I.simplify(in2))


def hidden_irreps(l_max, num_features, parity):
    out = []
    for l in range(l_max + 1):
        for p in ((1, -1) if parity else ((1,) if l % 2 == 0 else (-1,))):
            out.append((num_features, (l, p)))
    return out


def mlp(x, weights, alphas):
    for li, (W, a) in enumerate(zip(weights, alphas)):
        x = torch.mm(x, W * a)
        if li != len(weights) - 1:
            x = torch.nn.functional.silu(x)
    return x


# ------------------------------------------------------------------ the model
def energy(sd: Dict[str, torch.Tensor], cfg: dict, data: dict, model_dtype=torch.float32, tp_chunk: int = 0):
    """Total energy [1,1] f64 and per-atom energies from ``state_dict`` sd of NequIPEnergyModel.
    With ``data["edge_vectors"]`` given the edge geometry is taken from it (the ML-IAP branch,
    nequip/nn/utils.py:68-118 ``with_edge_vectors_`` keeps vectors that are already present)."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    pos, edge_index, types = data["pos"], data["edge_index"], data["atom_types"].view(-1)
    cell, shift = data.get("cell"), data.get("edge_cell_shift")
    if cell is None:
        shift = None
    l_max, nf = cfg["l_max"], cfg["num_features"]
    sh_ir = I.spherical_harmonics(l_max)
    if "edge_vectors" in data:
        vec = data["edge_vectors"]
        r = vec.square().sum(1, keepdim=True).sqrt()
        y = osh.spherical_harmonics(l_max, vec, normalize=True).to(model_dtype)
        emb = radial_embedding(r, cfg["r_max"], cfg["num_bessels"], float(cfg["polynomial_cutoff_p"]), model_dtype)
    else:
        _, y, emb = edge_embed(pos, edge_index, cell, shift, l_max, cfg["num_bessels"], cfg["r_max"],
                               float(cfg["polynomial_cutoff_p"]), model_dtype)
    node_attrs = sd["type_embed.weight"].to(model_dtype)[types]
    x = node_attrs
    prev = [(nf, (0, 1))]
    hid = hidden_irreps(l_max, nf, cfg["parity"])
    hiddens = [hid] * (cfg["num_layers"] - 1) + [[(nf, (0, 1))]]
    ann = cfg["avg_num_neighbors"]  # global, or one value per type (AvgNumNeighborsNorm, nequip/nn/norm.py:39-68)
    if isinstance(ann, (int, float)):
        norm = torch.tensor(1.0 / math.sqrt(ann), dtype=model_dtype)
    else:
        norm = torch.tensor([1.0 / math.sqrt(v) for v in ann], dtype=model_dtype)[types].view(-1, 1)
    depth = cfg["radial_mlp_depth"]
    for li, h in enumerate(hiddens):
        scalars = [(m, ir) for m, ir in h if ir[0] == 0 and tp_path_exists(prev, sh_ir, ir)]
        gated = [(m, ir) for m, ir in h if ir[0] > 0 and tp_path_exists(prev, sh_ir, ir)]
        gate_ir = (0, 1) if tp_path_exists(prev, sh_ir, (0, 1)) else (0, -1)
        gates = [(m, gate_ir) for m, _ in gated]
        conv_out = I.simplify(scalars + gates + gated)
        mid, ins = I.build_tp_instructions(prev, sh_ir, conv_out)
        pre = f"layers.{li}.conv."
        sc = None
        if li != 0:
            sc = fctp_scalar_attr(x, node_attrs, sd[pre + "sc.weight"].to(model_dtype), prev, nf, conv_out)
        x = linear(x, sd[pre + "linear_1.weight"].to(model_dtype), prev, prev)
        x = x * norm
        dims = [cfg["num_bessels"]] + depth * [cfg["radial_mlp_width"]] + [otp.weight_numel(prev, sh_ir, ins)]
        ws, alphas = [], []
        for q in range(depth + 1):
            ws.append(sd[pre + f"edge_mlp.mlp.{2 * q}.weight"].to(model_dtype))
            gain = 1.0 if q == 0 else math.sqrt(2)
            alphas.append(torch.tensor(gain / math.sqrt(dims[q]), dtype=model_dtype))
        w = mlp(emb, ws, alphas)
        x = otp.tp_scatter(x, y, w, edge_index[0], edge_index[1], prev, sh_ir, mid, ins, chunk=tp_chunk)
        x = linear(x, sd[pre + "linear_2.weight"].to(model_dtype), I.simplify(mid), conv_out)
        if sc is not None:
            x = x + sc
        x = gate(x, scalars, gates, gated)
        prev = scalars + [(m, (l, p * gate_ir[1])) for m, (l, p) in gated]
    wr = sd["readout.mlp.0.weight"].to(model_dtype)
    e_atom = torch.mm(x, wr * torch.tensor(1.0 / math.sqrt(wr.shape[0]), dtype=model_dtype)).to(torch.float64)
    if "scales" in sd and sd["scales"].numel():
        e_atom = e_atom * sd["scales"][types]
    if "shifts" in sd and sd["shifts"].numel():
        e_atom = e_atom + sd["shifts"][types]
    if data.get("batch") is not None:  # AtomwiseReduce per graph (nequip/nn/atomwise.py:92-113): [num_graphs, 1]
        batch = data["batch"].view(-1).long()
        ng = int(data["num_atoms"].numel()) if "num_atoms" in data else (int(batch.max()) + 1 if batch.numel() else 0)
        return torch.zeros((ng, 1), dtype=e_atom.dtype).index_add(0, batch, e_atom), e_atom
    return e_atom.sum(0, keepdim=True), e_atom


def energy_and_forces(sd, cfg, data, model_dtype=torch.float32, tp_chunk: int = 0):
    data = dict(data)
    pos = data["pos"].detach().clone().requires_grad_(True)
    data["pos"] = pos
    e_tot, e_atom = energy(sd, cfg, data, model_dtype, tp_chunk)
    (g,) = torch.autograd.grad([e_tot.sum()], [pos])
    return e_tot.detach(), e_atom.detach(), -g


def energy_forces_stress(sd, cfg, data, model_dtype=torch.float32, tp_chunk: int = 0):
    """ForceStressOutput (nequip/nn/grad_output.py:162-268): symmetric infinitesimal displacement applied to
    positions and cell, forces = -dE/dpos, virial_raw = dE/d(displacement), stress = virial_raw / |det cell|,
    virial = -virial_raw.  Returns (E, forces, stress [1,3,3], virial [1,3,3])."""
    data = dict(data)
    pos = data["pos"].detach().clone().requires_grad_(True)
    disp = torch.zeros(3, 3, dtype=pos.dtype, requires_grad=True)
    sym = 0.5 * (disp + disp.t())
    data["pos"] = pos + torch.sum(pos.view(-1, 3, 1) * sym, 1)
    cell = data["cell"].view(3, 3)
    data["cell"] = cell + torch.sum(cell.view(3, 3, 1) * sym, 1)
    e_tot, _ = energy(sd, cfg, data, model_dtype, tp_chunk)
    g, v = torch.autograd.grad([e_tot.sum()], [pos, disp])
    vol = torch.linalg.det(cell).abs()
    return e_tot.detach(), -g, (v / vol).view(1, 3, 3), (-v).view(1, 3, 3)


def edge_forces(sd, cfg, data, model_dtype=torch.float32):
    """The ML-IAP branch of ForceStressOutput (grad_output.py:270-296): dE/d(edge_vectors), no sign flip."""
    data = dict(data)
    vec = data["edge_vectors"].detach().clone().requires_grad_(True)
    data["edge_vectors"] = vec
    e_tot, _ = energy(sd, cfg, data, model_dtype)
    (g,) = torch.autograd.grad([e_tot.sum()], [vec])
    return e_tot.detach(), g
