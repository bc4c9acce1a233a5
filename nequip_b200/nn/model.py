"""Host-side (PyTorch) mirror of the NequIP energy model around the fused kernels.

This is *plumbing*: the module order, irreps bookkeeping, initialisation and
normalisation constants of the reference's model builder, so that the hot-path
kernels can be driven end to end (energy + forces) on identical
``AtomicDataDict``-shaped batches without e3nn/nequip installed.  Everything on the
per-edge hot path goes through ``nequip_b200.ops`` (CUDA); node-side dense algebra
uses torch matmul (cuBLAS, fp32, TF32 off).

Reference files mirrored (under /root/reference):
  model assembly          nequip/model/nequip_models.py:116-210 (NequIPGNNModel), :214-399 (Full...)
  ConvNetLayer            nequip/nn/convnetlayer.py:74-170
  InteractionBlock        nequip/nn/interaction_block.py:21-207
  ScalarMLPFunction       nequip/nn/mlp.py:80-195, ScalarLinearLayer :223-271
  AvgNumNeighborsNorm     nequip/nn/norm.py:7-68
  NodeTypeEmbed           nequip/nn/embedding/node.py:146-175
  PerTypeScaleShift       nequip/nn/atomwise.py:236-284;  AtomwiseReduce :92-113
  ForceStressOutput       nequip/nn/grad_output.py:107-298 (forces only)
  e3nn o3.Linear / FullyConnectedTensorProduct / nn.Gate semantics: SURVEY.md Appendix A.4
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import ops
from ..irreps import Irrep, Irreps, build_tp_instructions, tp_path_exists
from .tp_scatter import B200TensorProductScatter

# e3nn.math.normalize2mom constants: (E_{z~N(0,1)} act(z)^2)^(-1/2) estimated from
# torch.randn(1_000_000, generator=Generator("cpu").manual_seed(0), dtype=float64)
C_SILU = 1.6791767923989418
C_TANH = 1.5937334472592692

# AtomicDataDict keys used here (nequip/data/_keys.py:14-115)
POSITIONS_KEY = "pos"
EDGE_INDEX_KEY = "edge_index"
EDGE_CELL_SHIFT_KEY = "edge_cell_shift"
CELL_KEY = "cell"
ATOM_TYPE_KEY = "atom_types"
TOTAL_ENERGY_KEY = "total_energy"
PER_ATOM_ENERGY_KEY = "atomic_energy"
FORCE_KEY = "forces"
STRESS_KEY = "stress"
VIRIAL_KEY = "virial"
EDGE_VECTORS_KEY = "edge_vectors"
EDGE_FORCE_KEY = "edge_forces"
BATCH_KEY = "batch"
BATCH_PTR_KEY = "ptr"
NUM_NODES_KEY = "num_atoms"


# ---------------------------------------------------------------------------------------
# irreps bookkeeping of the conv stack
# ---------------------------------------------------------------------------------------
def hidden_irreps(l_max: int, num_features: int, parity: bool) -> Irreps:
    """``feature_irreps_hidden`` of NequIPGNNModel (nequip_models.py:176-187)."""
    items = []
    for l in range(l_max + 1):
        ps = (1, -1) if parity else ((1,) if l % 2 == 0 else (-1,))
        for p in ps:
            items.append((num_features, Irrep(l, p)))
    return Irreps(items)


def gate_irreps(prev: Irreps, edge_attr: Irreps, hidden: Irreps):
    """Irreps decisions of ConvNetLayer.__init__ (convnetlayer.py:74-114) for the gate nonlinearity.
    Returns (irreps_scalars, irreps_gates, irreps_gated, conv_irreps_out, layer_out)."""
    scalars = Irreps([(mul, ir) for mul, ir in hidden if ir.l == 0 and tp_path_exists(prev, edge_attr, ir)])
    gated = Irreps([(mul, ir) for mul, ir in hidden if ir.l > 0 and tp_path_exists(prev, edge_attr, ir)])
    gate_ir = Irrep(0, 1) if tp_path_exists(prev, edge_attr, Irrep(0, 1)) else Irrep(0, -1)
    gates = Irreps([(mul, gate_ir) for mul, _ in gated])
    conv_out = (scalars + gates + gated).simplify()
    # Gate.irreps_out = scalars + gated (with gates of even parity the gated irreps are unchanged)
    gated_out = Irreps([(mul, Irrep(ir.l, ir.p * gate_ir.p)) for mul, ir in gated])
    layer_out = scalars + gated_out
    return scalars, gates, gated, conv_out, layer_out


def layer_irreps(l_max: int, num_features: int, num_layers: int, parity: bool = True, type_embed_num_features=None):
    """[(feature_irreps_in, irreps_edge_attr, conv_irreps_out, (scalars, gates, gated))] per layer."""
    f0 = type_embed_num_features or num_features
    edge_attr = Irreps.spherical_harmonics(l_max)
    prev = Irreps([(f0, Irrep(0, 1))])
    hid = hidden_irreps(l_max, num_features, parity)
    hiddens = [hid] * (num_layers - 1) + [Irreps([(num_features, Irrep(0, 1))])]
    out = []
    for h in hiddens:
        scalars, gates, gated, conv_out, layer_out = gate_irreps(prev, edge_attr, h)
        out.append((prev, edge_attr, conv_out, (scalars, gates, gated)))
        prev = layer_out
    return out


# ---------------------------------------------------------------------------------------
# dense pieces (torch)
# ---------------------------------------------------------------------------------------
class ScalarLinearLayer(torch.nn.Module):
    """mlp.py:223-271: ``mm(input, weight * alpha)``, weight ~ U(-sqrt3, sqrt3)."""

    def __init__(self, in_features: int, out_features: int, alpha: float = 1.0):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("alpha", torch.tensor(alpha), persistent=False)
        self.weight = torch.nn.Parameter(torch.empty((in_features, out_features)))
        torch.nn.init.uniform_(self.weight, -math.sqrt(3), math.sqrt(3))

    def forward(self, x):
        return torch.mm(x, self.weight * self.alpha)


class ScalarMLPFunction(torch.nn.Module):
    """mlp.py:80-195 with ``bias=False, forward_weight_init=True, nonlinearity="silu"``."""

    def __init__(self, input_dim: int, output_dim: int, hidden_layers_depth: int = 0,
                 hidden_layers_width: Optional[int] = None, nonlinearity: Optional[str] = "silu"):
        super().__init__()
        dims = [input_dim] + hidden_layers_depth * [hidden_layers_width] + [output_dim]
        self.dims = dims
        layers: List[torch.nn.Module] = []
        nl = len(dims) - 1
        for layer, (h_in, h_out) in enumerate(zip(dims, dims[1:])):
            gain = 1.0 if nonlinearity is None or layer == 0 else math.sqrt(2)
            layers.append(ScalarLinearLayer(h_in, h_out, alpha=gain / math.sqrt(h_in)))
            if layer != nl - 1 and nonlinearity is not None:
                layers.append(torch.nn.SiLU())
        self.mlp = torch.nn.Sequential(*layers)

    def forward(self, x):
        return self.mlp(x)


class Linear(torch.nn.Module):
    """e3nn ``o3.Linear(irreps_in, irreps_out)`` (internal shared weights, no bias,
    path_normalization="element"): out_b = (1/sqrt(sum_a mul_a)) sum_a x_a W_ab over equal irreps."""

    def __init__(self, irreps_in, irreps_out, layout: str = "mul_ir"):
        super().__init__()
        self.layout = layout
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.instr: List[Tuple[int, int, int, float]] = []  # (i_in, i_out, weight offset, path weight)
        off = 0
        pairs = [
            (i, o)
            for i, (_, ir_i) in enumerate(self.irreps_in)
            for o, (_, ir_o) in enumerate(self.irreps_out)
            if ir_i == ir_o
        ]
        for (i, o) in pairs:
            fan = sum(self.irreps_in[i2][0] for (i2, o2) in pairs if o2 == o)
            self.instr.append((i, o, off, 1.0 / math.sqrt(fan)))
            off += self.irreps_in[i][0] * self.irreps_out[o][0]
        self.weight_numel = off
        self.weight = torch.nn.Parameter(torch.randn(off))
        self._in_sl, self._out_sl = self.irreps_in.slices(), self.irreps_out.slices()

    def forward(self, x):
        N = x.shape[0]
        outs: List[Optional[torch.Tensor]] = [None] * len(self.irreps_out)
        for (i, o, off, pw) in self.instr:
            mi, ir = self.irreps_in[i]
            mo = self.irreps_out[o][0]
            W = self.weight[off: off + mi * mo].view(mi, mo) * pw
            if self.layout == "ir_mul":
                # chunk = [d, mi] per node: one GEMM over all (node, component) rows
                r = torch.matmul(x[:, self._in_sl[i]].reshape(N * ir.dim, mi), W).reshape(N, ir.dim * mo)
            else:
                xi = x[:, self._in_sl[i]].reshape(N, mi, ir.dim)
                # [N, d, mi] @ [mi, mo] -> [N, d, mo] -> [N, mo, d]
                r = torch.matmul(xi.transpose(1, 2), W).transpose(1, 2).reshape(N, mo * ir.dim)
            outs[o] = r if outs[o] is None else outs[o] + r
        for o, (mo, ir) in enumerate(self.irreps_out):
            if outs[o] is None:
                outs[o] = x.new_zeros(N, mo * ir.dim)
        return torch.cat(outs, dim=1)


class SelfConnection(torch.nn.Module):
    """e3nn ``FullyConnectedTensorProduct(feature_irreps_in, F0 x 0e, feature_irreps_out)``
    (interaction_block.py:140-146): out_b[w,k] = (1/sqrt(sum_a mul_a F0)) sum_a sum_uv W_ab[u,v,w] x_a[u,k] attr[v]."""

    def __init__(self, irreps_in, num_attr: int, irreps_out, layout: str = "mul_ir"):
        super().__init__()
        self.layout = layout
        self.irreps_in, self.irreps_out, self.num_attr = Irreps(irreps_in), Irreps(irreps_out), num_attr
        pairs = [
            (i, o)
            for i, (_, ir_i) in enumerate(self.irreps_in)
            for o, (_, ir_o) in enumerate(self.irreps_out)
            if ir_i == ir_o
        ]
        self.instr: List[Tuple[int, int, int, float]] = []
        off = 0
        for (i, o) in pairs:
            fan = sum(self.irreps_in[i2][0] * num_attr for (i2, o2) in pairs if o2 == o)
            self.instr.append((i, o, off, 1.0 / math.sqrt(fan)))
            off += self.irreps_in[i][0] * num_attr * self.irreps_out[o][0]
        self.weight_numel = off
        self.weight = torch.nn.Parameter(torch.randn(off))
        self._in_sl = self.irreps_in.slices()

    def forward(self, x, node_attrs, types=None, type_table=None):
        """Generic form: ``node_attrs`` [N, F0].  When the attributes are a per-type table
        (``node_attrs == type_table[types]``, which is how NequIP builds them,
        nequip/nn/embedding/node.py:146-175) the bilinear form collapses to one GEMM per irrep
        pair with the per-type effective weights  Weff[t] = sum_v W[:, v, :] table[t, v]."""
        N = x.shape[0]
        outs: List[Optional[torch.Tensor]] = [None] * len(self.irreps_out)
        fast = types is not None and type_table is not None
        if fast:
            T = type_table.shape[0]
            onehot = torch.nn.functional.one_hot(types, T).to(x.dtype)  # [N, T]
        for (i, o, off, pw) in self.instr:
            mi, ir = self.irreps_in[i]
            mo = self.irreps_out[o][0]
            W = self.weight[off: off + mi * self.num_attr * mo].view(mi, self.num_attr, mo)
            if self.layout == "ir_mul":
                xk = x[:, self._in_sl[i]].reshape(N, ir.dim, mi)  # [N, d, mi]
                if fast:
                    weff = torch.einsum("uvw,tv->tuw", W, type_table).reshape(T * mi, mo) * pw
                    xe = (onehot.view(N, 1, T, 1) * xk.unsqueeze(2)).reshape(N * ir.dim, T * mi)
                    r = torch.matmul(xe, weff).reshape(N, ir.dim * mo)
                else:
                    r = pw * torch.einsum("uvw,nku,nv->nkw", W, xk, node_attrs).reshape(N, ir.dim * mo)
                outs[o] = r if outs[o] is None else outs[o] + r
                continue
            xi = x[:, self._in_sl[i]].reshape(N, mi, ir.dim)
            if fast:
                weff = torch.einsum("uvw,tv->tuw", W, type_table).reshape(T * mi, mo) * pw
                # [N, d, T*mi] @ [T*mi, mo]
                xe = (onehot.view(N, 1, T, 1) * xi.transpose(1, 2).unsqueeze(2)).reshape(N, ir.dim, T * mi)
                r = torch.matmul(xe, weff).transpose(1, 2).reshape(N, mo * ir.dim)
            else:
                r = pw * torch.einsum("uvw,nuk,nv->nwk", W, xi, node_attrs).reshape(N, mo * ir.dim)
            outs[o] = r if outs[o] is None else outs[o] + r
        for o, (mo, ir) in enumerate(self.irreps_out):
            if outs[o] is None:
                outs[o] = x.new_zeros(N, mo * ir.dim)
        return torch.cat(outs, dim=1)


class Gate(torch.nn.Module):
    """e3nn ``nn.Gate`` with normalize2mom'd SiLU (even) / tanh (odd) (convnetlayer.py:42-56,104-112)."""

    def __init__(self, irreps_scalars, irreps_gates, irreps_gated, layout: str = "mul_ir"):
        super().__init__()
        self.layout = layout
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = (
            Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated))
        self.irreps_in = self.irreps_scalars + self.irreps_gates + self.irreps_gated
        gp = self.irreps_gates[0][1].p if len(self.irreps_gates) else 1
        self.irreps_out = self.irreps_scalars + Irreps([(m, Irrep(ir.l, ir.p * gp)) for m, ir in self.irreps_gated])
        self.use_fused = True  # CUDA: fused kernels; the torch formulation below is the readable definition
        self._tabs = None

    @staticmethod
    def _act(x, p: int):
        return torch.nn.functional.silu(x) * C_SILU if p == 1 else torch.tanh(x) * C_TANH

    def forward(self, x):
        if x.is_cuda and self.use_fused:
            # one kernel per direction (nqb_gate_fwd/bwd) instead of ~30 strided elementwise ops
            key = (x.device, self.layout)
            if self._tabs is None or self._tabs[0] != key:
                self._tabs = (key, ops.GateTables(self.irreps_scalars, self.irreps_gates, self.irreps_gated,
                                                  self.layout, x.device))
            return ops.gate(x, self._tabs[1])
        N = x.shape[0]
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        parts = []
        off = 0
        for mul, ir in self.irreps_scalars:
            parts.append(self._act(x[:, off: off + mul], ir.p))
            off += mul
        if len(self.irreps_gated):
            gates = []
            goff = ns
            for mul, ir in self.irreps_gates:
                gates.append(self._act(x[:, goff: goff + mul], ir.p))
                goff += mul
            gates = torch.cat(gates, dim=1)
            off = ns + ng
            g0 = 0
            for mul, ir in self.irreps_gated:
                if self.layout == "ir_mul":
                    ch = x[:, off: off + mul * ir.dim].reshape(N, ir.dim, mul)
                    parts.append((ch * gates[:, g0: g0 + mul].unsqueeze(1)).reshape(N, mul * ir.dim))
                else:
                    ch = x[:, off: off + mul * ir.dim].reshape(N, mul, ir.dim)
                    parts.append((ch * gates[:, g0: g0 + mul].unsqueeze(-1)).reshape(N, mul * ir.dim))
                off += mul * ir.dim
                g0 += mul
        return torch.cat(parts, dim=1)


# ---------------------------------------------------------------------------------------
# graph modules
# ---------------------------------------------------------------------------------------
class InteractionBlock(torch.nn.Module):
    """interaction_block.py:21-207 (no ghost exchange here; see nequip_b200.parallel for the sharded path)."""

    def __init__(self, feature_irreps_in, irreps_edge_attr, feature_irreps_out, num_edge_embed: int,
                 num_node_attrs: int, radial_mlp_depth: int, radial_mlp_width: int, use_sc: bool,
                 avg_num_neighbors: float, is_first_layer: bool = False, layout: str = "mul_ir"):
        super().__init__()
        self.is_first_layer = is_first_layer
        self.layout = layout
        fin, fe, fout = Irreps(feature_irreps_in), Irreps(irreps_edge_attr), Irreps(feature_irreps_out)
        self.feature_irreps_in, self.irreps_edge_attr, self.feature_irreps_out = fin, fe, fout
        # AvgNumNeighborsNorm (nequip/nn/norm.py:7-68): a global value, or one per atom type (a sequence in
        # type order / the reference's dict after ordering by type_names); norm_const = 1 / sqrt(avg_num_neighbors)
        ann = [float(avg_num_neighbors)] if isinstance(avg_num_neighbors, (int, float)) else [float(v) for v in avg_num_neighbors]
        self.register_buffer("norm_const", torch.tensor([1.0 / math.sqrt(v) for v in ann]).reshape(-1, 1), persistent=False)
        self.norm_shortcut = len(ann) == 1
        self.linear_1 = Linear(fin, fin, layout)
        irreps_mid, instructions = build_tp_instructions(fin, fe, fout)
        self.irreps_mid, self.instructions = irreps_mid, instructions
        self.tp_scatter = B200TensorProductScatter(fin, fe, irreps_mid, instructions, layout=layout)
        self.edge_mlp = ScalarMLPFunction(num_edge_embed, self.tp_scatter.weight_numel,
                                          hidden_layers_depth=radial_mlp_depth,
                                          hidden_layers_width=radial_mlp_width, nonlinearity="silu")
        self.linear_2 = Linear(irreps_mid.simplify(), fout, layout)
        self.sc = SelfConnection(fin, num_node_attrs, fout, layout) if use_sc else None
        self.use_tensor_cores = True
        self.use_fused_radial_tp = "auto"  # SURVEY 8f-1 kernel (mul % 32 == 0, K <= 128): True / False / "auto" (timed once)
        self._fused_choice = None
        self.strict_fast_path = False
        self._tc_cache = None

    def _edge_weights(self, edge_embedding):
        """Radial MLP, plain torch.mm formulation of the reference (mlp.py:262-268) -- the path for trainable
        weights, float64 and unusual shapes; frozen float32 ir_mul models use ``_tensor_core_blocks``."""
        return self.edge_mlp(edge_embedding)

    def _use_fused(self, tc, edge_embedding, x, edge_attrs, edge_index) -> bool:
        """``use_fused_radial_tp``: True / False, or "auto" (default) = time the fused kernel against the unfused pair
        (grouped GEMM + TP kernel) ONCE per layer on the first real call and keep the faster one.  The fused kernel
        never materialises the [E, W] weights in the forward pass, but its path-parallel decomposition gives up the
        sharing of the x_i Y_j products between paths: which one wins depends on the signature (measurements in
        DESIGN.md section 4.7)."""
        mode = self.use_fused_radial_tp
        if mode is True or mode is False:
            return mode
        if self._fused_choice is None:
            if torch.cuda.is_current_stream_capturing() or tc["mlp"] is None:
                return False
            with torch.no_grad():
                def t(fn):
                    fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(2):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / 2

                xd, yd, ed = x.detach(), edge_attrs.detach(), edge_embedding.detach()
                tf = t(lambda: tc["fused"](ed, xd, yd, edge_index[0], edge_index[1]))
                tu = t(lambda: self.tp_scatter(x=xd, edge_attr=yd, edge_weight=tc["mlp"](ed), edge_dst=edge_index[0],
                                               edge_src=edge_index[1]))
            self._fused_choice = bool(tf < tu)
            self.fused_timing_ms = {"fused": tf, "unfused": tu}
        return self._fused_choice

    def _note_fallback(self, reason: str):
        """The library (torch.matmul / cuBLAS) formulation is about to run instead of the tcgen05 blocks: say so
        once per block and reason, or raise when the model was built with ``strict_fast_path=True`` (bench.py,
        smoke(): a silent library fallback would be timed as if it were the product)."""
        if getattr(self, "strict_fast_path", False):
            raise RuntimeError(f"nequip_b200 InteractionBlock: tensor-core fast path unavailable ({reason}) and "
                               "strict_fast_path=True")
        seen = self.__dict__.setdefault("_fallback_seen", set())
        if reason not in seen:
            seen.add(reason)
            import warnings

            warnings.warn(f"nequip_b200 InteractionBlock: dense blocks run through torch.matmul (cuBLAS), not the "
                          f"tcgen05 kernels: {reason}", RuntimeWarning, stacklevel=3)

    def forward(self, x, node_attrs, edge_attrs, edge_embedding, edge_index, types=None, type_table=None,
                n_own: Optional[int] = None, halo=None):
        """``n_own``/``halo``: sharded frames (owned atoms first, then ghosts).  As in the reference
        (interaction_block.py:159-199) the first layer sees the type embedding of owned + ghost atoms;
        later layers work on owned rows, refresh the ghosts through ``halo`` right before the
        TP+scatter and truncate to the owned rows right after it."""
        if n_own is not None and not self.is_first_layer:
            x = x[:n_own]
            node_attrs = node_attrs[:n_own]
            types = None if types is None else types[:n_own]
        tc = self._tensor_core_blocks(x, types, type_table)
        if tc is not None:
            # inference fast path: every dense block is one grouped 3xTF32 tcgen05 GEMM launch
            x_in = x
            # 1/sqrt(avg_num_neighbors): folded into the prepared weights (global) or a per-atom row scale (per type)
            x = tc["lin1"](x) if self.norm_shortcut else tc["lin1"](x, self.norm_const.view(-1)[types].view(1, -1).contiguous())
            if halo is not None and not self.is_first_layer:
                x = halo(x)
            if tc["fused"] is not None and self._use_fused(tc, edge_embedding, x, edge_attrs, edge_index):
                # one kernel: last radial layer (tcgen05, weights resident in tensor memory) -> TP -> scatter
                y = tc["fused"](edge_embedding, x, edge_attrs, edge_index[0], edge_index[1])
                if y is not None:
                    x = y
                    if n_own is not None:
                        x = x[:n_own]
                    x = tc["lin2"](x)
                    if tc["sc"] is not None:
                        x = tc["sc"](x_in, types, x)
                    return x
            if tc["mlp"] is not None:
                w = tc["mlp"](edge_embedding)
            else:
                self._note_fallback("radial MLP shape not supported by the grouped GEMM (needs one hidden layer, "
                                    "widths multiples of 4)")
                w = self._edge_weights(edge_embedding)
            x = self.tp_scatter(x=x, edge_attr=edge_attrs, edge_weight=w, edge_dst=edge_index[0], edge_src=edge_index[1])
            if n_own is not None:
                x = x[:n_own]
            x = tc["lin2"](x)
            if tc["sc"] is not None:
                x = tc["sc"](x_in, types, x)  # accumulates the self-connection onto linear_2's output
            return x
        sc = self.sc(x, node_attrs, types, type_table) if self.sc is not None else None
        x = self.linear_1(x)
        x = x * (self.norm_const.view(()) if self.norm_shortcut else self.norm_const[types])
        if halo is not None and not self.is_first_layer:
            x = halo(x)
        w = self._edge_weights(edge_embedding)
        x = self.tp_scatter(x=x, edge_attr=edge_attrs, edge_weight=w, edge_dst=edge_index[0], edge_src=edge_index[1])
        if n_own is not None:
            x = x[:n_own]
        x = self.linear_2(x)
        if sc is not None:
            x = x + sc
        return x

    def _tensor_core_blocks(self, x, types, type_table):
        """Lazily prepared tensor-core versions of linear_1 / radial MLP / linear_2 / self-connection
        (nequip_b200/nn/dense.py); None when not applicable (training, float64, mul_ir, odd multiplicities)."""
        from . import dense

        if not self.use_tensor_cores or not x.is_cuda:
            return None
        if x.dtype != torch.float32 or self.layout != "ir_mul":
            self._note_fallback(f"dtype {x.dtype} / layout {self.layout} (needs float32, ir_mul)")
            return None
        if any(p.requires_grad for p in self.parameters()) or (type_table is not None and type_table.requires_grad):
            self._note_fallback("parameters require grad (training); freeze them for the inference path")
            return None
        if (self.sc is not None or not self.norm_shortcut) and (types is None or type_table is None):
            self._note_fallback("self-connection / per-type normalisation need atom types + type table")
            return None
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (
            (type_table.data_ptr(), type_table._version) if type_table is not None else ())
        if self._tc_cache is not None and self._tc_cache[0] == key:
            if self._tc_cache[1] is None:
                self._note_fallback("a multiplicity is not a multiple of 4")
            return self._tc_cache[1]
        ok = dense.IrrepsLinearGemm.supported(self.linear_1) and dense.IrrepsLinearGemm.supported(self.linear_2)
        if self.sc is not None:
            ok = ok and dense.SelfConnectionGemm.supported(self.sc)
        if not ok:
            self._note_fallback("a multiplicity is not a multiple of 4")
            self._tc_cache = (key, None)
            return None
        dev = x.device
        lins = [m for m in self.edge_mlp.mlp if isinstance(m, ScalarLinearLayer)]
        mlp = None
        if len(lins) == 2 and dense.RadialMLPGemm.supported(lins[0], lins[1], x.dtype):
            mlp = dense.RadialMLPGemm(lins[0], lins[1], dev)
        fused = None
        if len(lins) == 2 and dense.FusedRadialTP.supported(lins[0], lins[1], self.tp_scatter._plan, x.dtype):
            fused = dense.FusedRadialTP(lins[0], lins[1], self.tp_scatter._plan, dev)
        blocks = dict(
            fused=fused,
            lin1=(dense.IrrepsLinearGemm(self.linear_1, dev, extra_scale=float(self.norm_const.view(-1)[0]))
                  if self.norm_shortcut else dense.IrrepsLinearGemm(self.linear_1, dev, row_scaled=True)),
            lin2=dense.IrrepsLinearGemm(self.linear_2, dev),
            sc=dense.SelfConnectionGemm(self.sc, type_table, dev) if self.sc is not None else None,
            mlp=mlp,
        )
        self._tc_cache = (key, blocks)
        return blocks


class ConvNetLayer(torch.nn.Module):
    """convnetlayer.py:26-170 (gate nonlinearity, no resnet)."""

    def __init__(self, prev: Irreps, edge_attr: Irreps, hidden: Irreps, **conv_kwargs):
        super().__init__()
        scalars, gates, gated, conv_out, layer_out = gate_irreps(prev, edge_attr, hidden)
        self.equivariant_nonlin = Gate(scalars, gates, gated, conv_kwargs.get("layout", "mul_ir"))
        self.conv = InteractionBlock(prev, edge_attr, conv_out, **conv_kwargs)
        self.irreps_out = layer_out

    def forward(self, x, node_attrs, edge_attrs, edge_embedding, edge_index, types=None, type_table=None,
                n_own=None, halo=None):
        x = self.conv(x, node_attrs, edge_attrs, edge_embedding, edge_index, types, type_table, n_own, halo)
        return self.equivariant_nonlin(x)


class NequIPEnergyModel(torch.nn.Module):
    """``NequIPGNNModel`` (nequip_models.py:116-210) wrapped in the force part of
    ``ForceStressOutput`` (grad_output.py:215-232).  ``forward(data) -> data`` on an
    AtomicDataDict-shaped dict: needs ``pos`` [N,3] f64, ``edge_index`` [2,E] i64,
    ``atom_types`` [N] i64 and, for periodic systems, ``cell`` [3,3] + ``edge_cell_shift`` [E,3]."""

    def __init__(self, *, r_max: float, type_names: Sequence[str], num_layers: int = 4, l_max: int = 1,
                 parity: bool = True, num_features: int = 32, radial_mlp_depth: int = 1,
                 radial_mlp_width: int = 128, num_bessels: int = 8, polynomial_cutoff_p: float = 6.0,
                 avg_num_neighbors: float = 1.0, per_type_energy_scales: Optional[Sequence[float]] = None,
                 per_type_energy_shifts: Optional[Sequence[float]] = None, model_dtype=torch.float32,
                 seed: int = 123, node_layout: str = "ir_mul", strict_fast_path: bool = False):
        super().__init__()
        self.r_max, self.l_max, self.num_bessels, self.poly_p = float(r_max), l_max, num_bessels, float(polynomial_cutoff_p)
        self.model_dtype = model_dtype
        self.node_layout = node_layout  # internal layout of node features between the kernels
        self.config = dict(r_max=r_max, type_names=list(type_names), num_layers=num_layers, l_max=l_max, parity=parity,
                           num_features=num_features, radial_mlp_depth=radial_mlp_depth,
                           radial_mlp_width=radial_mlp_width, num_bessels=num_bessels,
                           polynomial_cutoff_p=polynomial_cutoff_p, avg_num_neighbors=avg_num_neighbors)
        prev_default = torch.get_default_dtype()
        torch.set_default_dtype(model_dtype)
        try:
            torch.manual_seed(seed)  # model_builder seeds before construction (model/utils.py:104-230)
            ntypes = len(type_names)
            if isinstance(avg_num_neighbors, dict):  # per type, keyed by type name (nequip/nn/norm.py:28-31)
                if set(avg_num_neighbors) != set(type_names):
                    raise ValueError("avg_num_neighbors: keys must be the type names")
                avg_num_neighbors = [float(avg_num_neighbors[k]) for k in type_names]
            elif not isinstance(avg_num_neighbors, (int, float)):
                avg_num_neighbors = [float(v) for v in avg_num_neighbors]
                if len(avg_num_neighbors) not in (1, ntypes):
                    raise ValueError(f"avg_num_neighbors: expected a scalar or {ntypes} values")
            self.config["avg_num_neighbors"] = avg_num_neighbors
            self.type_embed = torch.nn.Embedding(ntypes, num_features)
            edge_attr = Irreps.spherical_harmonics(l_max)
            prev = Irreps([(num_features, Irrep(0, 1))])
            hid = hidden_irreps(l_max, num_features, parity)
            hiddens = [hid] * (num_layers - 1) + [Irreps([(num_features, Irrep(0, 1))])]
            layers = []
            for li, h in enumerate(hiddens):
                layer = ConvNetLayer(prev, edge_attr, h, num_edge_embed=num_bessels, num_node_attrs=num_features,
                                     radial_mlp_depth=radial_mlp_depth, radial_mlp_width=radial_mlp_width,
                                     use_sc=(li != 0), avg_num_neighbors=avg_num_neighbors,
                                     is_first_layer=(li == 0), layout=node_layout)
                layers.append(layer)
                prev = layer.irreps_out
            self.layers = torch.nn.ModuleList(layers)
            self.readout = ScalarMLPFunction(prev.dim, 1, hidden_layers_depth=0)
        finally:
            torch.set_default_dtype(prev_default)
        # PerTypeScaleShift (atomwise.py:236-284): a float or a one-element list applies to every type
        # (the reference's scales_shortcut / shifts_shortcut); otherwise one value per type
        def table(v, what):
            if v is None:
                return torch.empty(0, dtype=torch.float64)
            t = torch.as_tensor(v, dtype=torch.float64).reshape(-1)
            if t.numel() == 1:
                t = t.expand(ntypes).clone()
            if t.numel() != ntypes:
                raise ValueError(f"{what}: expected a scalar or {ntypes} values (one per type), got {t.numel()}")
            return t.reshape(-1, 1)

        self.register_buffer("scales", table(per_type_energy_scales, "per_type_energy_scales"))
        self.register_buffer("shifts", table(per_type_energy_shifts, "per_type_energy_shifts"))
        self.set_strict_fast_path(strict_fast_path)

    def set_strict_fast_path(self, on: bool = True):
        """Raise instead of warning when an interaction block cannot use the tcgen05 dense blocks."""
        for layer in self.layers:
            layer.conv.strict_fast_path = bool(on)
        return self

    @staticmethod
    def _reduce_energy(e_atom: torch.Tensor, data: Dict[str, torch.Tensor]) -> torch.Tensor:
        """AtomwiseReduce (atomwise.py:92-113): per-graph sum -> [num_graphs, 1]; one frame without ``batch``."""
        batch = data.get(BATCH_KEY)
        if batch is None:
            return e_atom.sum(dim=0, keepdim=True)
        if NUM_NODES_KEY in data:
            ng = int(data[NUM_NODES_KEY].numel())
        elif BATCH_PTR_KEY in data:
            ng = int(data[BATCH_PTR_KEY].numel()) - 1
        else:
            ng = int(batch.max().item()) + 1 if batch.numel() else 0
        out = torch.zeros((ng, 1), dtype=e_atom.dtype, device=e_atom.device)
        return out.index_add_(0, batch.view(-1).long(), e_atom)

    # the energy part (SequentialGraphNetwork order of nequip_models.py:288-399)
    def energy(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        pos = data[POSITIONS_KEY]
        edge_index = data[EDGE_INDEX_KEY]
        types = data[ATOM_TYPE_KEY].view(-1)
        node_attrs = self.type_embed(types)
        x = node_attrs
        shift, cell = data.get(EDGE_CELL_SHIFT_KEY), data.get(CELL_KEY)
        if cell is None:
            shift = None
        pre = (2 * math.pi) / (self.r_max * self.r_max)
        if EDGE_VECTORS_KEY in data:
            # the caller (LAMMPS ML-IAP) supplies the edge vectors: with_edge_vectors_ keeps them (nn/utils.py:68-118)
            edge_attrs, edge_embedding = ops.edge_embed_from_vectors(
                data[EDGE_VECTORS_KEY], lmax=self.l_max, num_bessel=self.num_bessels, r_max=self.r_max,
                poly_p=self.poly_p, prefactor=pre, out_dtype=self.model_dtype)
        else:
            _vec, edge_attrs, edge_embedding = ops.edge_embed(
                pos, edge_index, shift, cell, lmax=self.l_max, num_bessel=self.num_bessels, r_max=self.r_max,
                poly_p=self.poly_p, prefactor=pre, out_dtype=self.model_dtype, edge_grad_sink=data.get("_edge_grad_sink"))
        for layer in self.layers:
            x = layer(x, node_attrs, edge_attrs, edge_embedding, edge_index, types, self.type_embed.weight)
        e_atom = self.readout(x).to(torch.float64)
        if self.scales.numel():
            e_atom = e_atom * self.scales[types]
        if self.shifts.numel():
            e_atom = e_atom + self.shifts[types]
        data[PER_ATOM_ENERGY_KEY] = e_atom
        data[TOTAL_ENERGY_KEY] = self._reduce_energy(e_atom, data)
        return data

    def energy_owned(self, data: Dict[str, torch.Tensor], n_own: int, halo) -> torch.Tensor:
        """Per-atom energies [n_own, 1] f64 of the OWNED atoms of a sharded frame (``data`` holds owned
        atoms first, then ghosts; every edge's destination is owned) -- see nequip_b200/parallel.py."""
        pos, edge_index = data[POSITIONS_KEY], data[EDGE_INDEX_KEY]
        types = data[ATOM_TYPE_KEY].view(-1)
        node_attrs = self.type_embed(types)
        x = node_attrs
        shift, cell = data.get(EDGE_CELL_SHIFT_KEY), data.get(CELL_KEY)
        if cell is None:
            shift = None
        _vec, edge_attrs, edge_embedding = ops.edge_embed(
            pos, edge_index, shift, cell, lmax=self.l_max, num_bessel=self.num_bessels, r_max=self.r_max,
            poly_p=self.poly_p, prefactor=(2 * math.pi) / (self.r_max * self.r_max), out_dtype=self.model_dtype)
        for layer in self.layers:
            x = layer(x, node_attrs, edge_attrs, edge_embedding, edge_index, types, self.type_embed.weight, n_own, halo)
        e_atom = self.readout(x).to(torch.float64)
        t_own = types[:n_own]
        if self.scales.numel():
            e_atom = e_atom * self.scales[t_own]
        if self.shifts.numel():
            e_atom = e_atom + self.shifts[t_own]
        return e_atom

    def forward(self, data: Dict[str, torch.Tensor], compute_forces: bool = True,
                compute_stress: bool = False) -> Dict[str, torch.Tensor]:
        """``ForceStressOutput.forward`` (nequip/nn/grad_output.py:107-298):

        * positions given: ``forces = -dE/dpos``; with ``compute_stress`` (needs ``cell``) also
          ``stress = (1/|det cell|) dE/d(eps)`` and ``virial = -dE/d(eps)`` ([1,3,3]) for the symmetric strain
          ``eps`` applied to positions and cell.  The cell/strain gradient is not taken through a displaced
          copy of the inputs: every edge vector transforms as ``r -> r (1 + eps)``, so
          ``dE/d(eps) = sym( sum_e r_e (x) dE/dr_e )`` and the per-edge gradients are a by-product of the
          edge-embedding backward kernel;
        * ``edge_vectors`` given (LAMMPS ML-IAP): ``edge_forces = dE/d(edge_vectors)``, no sign flip (:270-296).
        """
        data = dict(data)
        if not compute_forces:
            return self.energy(data)
        if EDGE_VECTORS_KEY in data:
            with torch.enable_grad():
                vec = data[EDGE_VECTORS_KEY].detach().double().requires_grad_(True)
                data[EDGE_VECTORS_KEY] = vec
                data = self.energy(data)
                (g,) = torch.autograd.grad([data[TOTAL_ENERGY_KEY].sum()], [vec])
            data[EDGE_FORCE_KEY] = g
            data[EDGE_VECTORS_KEY] = vec.detach()
            data[TOTAL_ENERGY_KEY] = data[TOTAL_ENERGY_KEY].detach()
            data[PER_ATOM_ENERGY_KEY] = data[PER_ATOM_ENERGY_KEY].detach()
            return data
        if compute_stress and data.get(CELL_KEY) is None:
            raise ValueError("compute_stress needs a cell")
        pos = data[POSITIONS_KEY]
        sink = {} if compute_stress else None
        with torch.enable_grad():
            pos = pos.detach().requires_grad_(True)
            data[POSITIONS_KEY] = pos
            if sink is not None:
                data["_edge_grad_sink"] = sink
            data = self.energy(data)
            (g,) = torch.autograd.grad([data[TOTAL_ENERGY_KEY].sum()], [pos])
        data.pop("_edge_grad_sink", None)
        data[FORCE_KEY] = torch.neg(g)
        if sink is not None:
            v = torch.einsum("ea,eb->ab", sink["edge_vectors"], sink["edge_vector_grad"])
            v = 0.5 * (v + v.t())
            vol = torch.linalg.det(data[CELL_KEY].double().view(3, 3)).abs()
            data[STRESS_KEY] = (v / vol).view(1, 3, 3)
            data[VIRIAL_KEY] = torch.neg(v).view(1, 3, 3)
        data[POSITIONS_KEY] = pos.detach()
        data[TOTAL_ENERGY_KEY] = data[TOTAL_ENERGY_KEY].detach()
        data[PER_ATOM_ENERGY_KEY] = data[PER_ATOM_ENERGY_KEY].detach()
        return data
