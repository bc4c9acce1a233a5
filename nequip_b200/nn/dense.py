"""Dense blocks of the interaction layer on the tensor cores (inference: frozen weights, float32,
channel-contiguous ``ir_mul`` node layout).

Each block is ONE launch of the grouped 3xTF32 GEMM (``nequip_b200/csrc/nqb_gemm.cu``) per
direction; the weights are prepared (scaled, split hi/lo, tiled) once.  Reference ops:

* ``RadialMLPGemm``      ScalarMLPFunction, depth 1            nequip/nn/mlp.py:80-195, 262-268
* ``IrrepsLinearGemm``   e3nn o3.Linear (linear_1, linear_2)   nequip/nn/interaction_block.py:82-87,129-138
* ``SelfConnectionGemm`` e3nn FullyConnectedTensorProduct(x, node_attrs) with node_attrs =
                         type_embed[atom_types]                nequip/nn/interaction_block.py:140-146,175

In ir_mul every (chunk pair, irrep component) is a strided GEMM over the atoms:
``out[:, oo + i*mo : oo + (i+1)*mo] (+)= x[:, io + i*mi : io + (i+1)*mi] @ W``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import ops
from ..irreps import Irreps


def _aligned(*vals) -> bool:
    return all(v % 4 == 0 for v in vals)


# ---------------------------------------------------------------------------------------
class _GemmLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd: ops.GroupedGemm, bwd: ops.GroupedGemm, d_out: int, zero_out: bool, d_in: int, zero_in: bool,
                rowscale, out_init):
        M = x.shape[0]
        if out_init is not None:
            out = out_init  # accumulate into an existing tensor (self-connection added onto linear_2's output)
        else:
            out = (torch.zeros if zero_out else torch.empty)((M, d_out), dtype=x.dtype, device=x.device)
        fwd.run(x, out, M, rowscale)
        ctx.bwd, ctx.d_in, ctx.zero_in, ctx.rowscale = bwd, d_in, zero_in, rowscale
        ctx.has_init = out_init is not None
        if out_init is not None:
            ctx.mark_dirty(out_init)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        gout = gout.contiguous()
        M = gout.shape[0]
        gx = (torch.zeros if ctx.zero_in else torch.empty)((M, ctx.d_in), dtype=gout.dtype, device=gout.device)
        ctx.bwd.run(gout, gx, M, ctx.rowscale)
        return gx, None, None, None, None, None, None, None, (gout if ctx.has_init else None)


class IrrepsLinearGemm:
    """``o3.Linear`` in ir_mul layout from a ``nequip_b200.nn.model.Linear`` module's weights."""

    def __init__(self, lin, device, extra_scale: float = 1.0, row_scaled: bool = False):
        """``row_scaled``: every problem multiplies its output rows by row 0 of the ``rowscale`` matrix given at call
        time (the per-atom-type AvgNumNeighborsNorm factor, nequip/nn/norm.py:48-68)."""
        fin, fout = lin.irreps_in, lin.irreps_out
        rs = 0 if row_scaled else -1
        self.row_scaled = row_scaled
        in_off, out_off = fin.offsets(), fout.offsets()
        self.d_in, self.d_out = fin.dim, fout.dim
        fwd: List[ops.GemmProblem] = []
        bwd: List[ops.GemmProblem] = []
        # problems of one launch run concurrently: a target written by more than one problem is
        # zero-initialised and every writer adds with red.global.add; single writers store plainly
        n_out = {o: sum(1 for (_, o2, _, _) in lin.instr if o2 == o) for o in range(len(fout))}
        n_in = {i: sum(1 for (i2, _, _, _) in lin.instr if i2 == i) for i in range(len(fin))}
        for (i, o, off, pw) in lin.instr:
            mi, ir = fin[i]
            mo = fout[o][0]
            W = lin.weight.detach()[off: off + mi * mo].view(mi, mo)
            for c in range(ir.dim):
                a_off, c_off = in_off[i] + c * mi, out_off[o] + c * mo
                fwd.append(ops.GemmProblem(a_off, self.d_in, c_off, self.d_out, W, scale=pw * extra_scale,
                                           atomic=n_out[o] > 1, rs_off=rs))
                bwd.append(ops.GemmProblem(c_off, self.d_out, a_off, self.d_in, W, scale=pw * extra_scale, transposed=True,
                                           atomic=n_in[i] > 1, rs_off=rs))
        self.zero_out = any(n != 1 for n in n_out.values())
        self.zero_in = any(n != 1 for n in n_in.values())
        self.fwd = ops.GroupedGemm(fwd, device)
        self.bwd = ops.GroupedGemm(bwd, device)

    @staticmethod
    def supported(lin) -> bool:
        muls = [m for m, _ in lin.irreps_in] + [m for m, _ in lin.irreps_out]
        return lin.layout == "ir_mul" and all(m % 4 == 0 for m in muls) and lin.weight.dtype == torch.float32

    def __call__(self, x, rowscale=None):
        if self.row_scaled != (rowscale is not None):
            raise ValueError("IrrepsLinearGemm: rowscale must be given exactly when built with row_scaled=True")
        return _GemmLinearFn.apply(x.contiguous(), self.fwd, self.bwd, self.d_out, self.zero_out, self.d_in, self.zero_in,
                                   rowscale, None)


class SelfConnectionGemm:
    """FCTP(x, type_embed[types]) as per-type effective-weight GEMMs with a one-hot row scale; the result
    is ACCUMULATED onto ``base`` (the linear_2 output), i.e. ``x = linear_2(x) + sc`` in one pass."""

    def __init__(self, sc, type_table: torch.Tensor, device):
        fin, fout = sc.irreps_in, sc.irreps_out
        in_off, out_off = fin.offsets(), fout.offsets()
        self.d_in, self.d_out = fin.dim, fout.dim
        self.T = type_table.shape[0]
        fwd: List[ops.GemmProblem] = []
        bwd: List[ops.GemmProblem] = []
        tt = type_table.detach()
        for (i, o, off, pw) in sc.instr:
            mi, ir = fin[i]
            mo = fout[o][0]
            W = sc.weight.detach()[off: off + mi * sc.num_attr * mo].view(mi, sc.num_attr, mo)
            weff = torch.einsum("uvw,tv->tuw", W, tt)  # [T, mi, mo]
            for t in range(self.T):
                for c in range(ir.dim):
                    a_off, c_off = in_off[i] + c * mi, out_off[o] + c * mo
                    # row scale = row t of one-hot^T [T, M]
                    # forward: each atom row belongs to exactly one type -> the T problems of one (pair,
                    # component) write disjoint rows; read-modify-write onto linear_2's output is race free
                    # as long as only ONE in-chunk feeds an out chunk, otherwise add atomically
                    multi_o = sum(1 for (_, o2, _, _) in sc.instr if o2 == o) > 1
                    fwd.append(ops.GemmProblem(a_off, self.d_in, c_off, self.d_out, weff[t].contiguous(), scale=pw,
                                               accumulate=not multi_o, atomic=multi_o, rs_off=t, skip_zero_rows=True))
                    multi_i = sum(1 for (i2, _, _, _) in sc.instr if i2 == i) > 1
                    bwd.append(ops.GemmProblem(c_off, self.d_out, a_off, self.d_in, weff[t].contiguous(), scale=pw,
                                               transposed=True, accumulate=not multi_i, atomic=multi_i, rs_off=t,
                                               skip_zero_rows=True))
        self.zero_in = True  # backward accumulates (row-masked) into a zero-initialised gradient
        self.fwd = ops.GroupedGemm(fwd, device)
        self.bwd = ops.GroupedGemm(bwd, device)

    @staticmethod
    def supported(sc) -> bool:
        muls = [m for m, _ in sc.irreps_in] + [m for m, _ in sc.irreps_out]
        return sc.layout == "ir_mul" and all(m % 4 == 0 for m in muls) and sc.weight.dtype == torch.float32

    def __call__(self, x, types, base):
        onehot_t = torch.nn.functional.one_hot(types, self.T).to(x.dtype).t().contiguous()  # [T, M]
        return _GemmLinearFn.apply(x.contiguous(), self.fwd, self.bwd, self.d_out, False, self.d_in, self.zero_in,
                                   onehot_t, base)


# ---------------------------------------------------------------------------------------
class _RadialMLPGemmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, w1s, fwd: ops.GroupedGemm, bwd: ops.GroupedGemm, W: int):
        E, hid = emb.shape[0], w1s.shape[1]
        fast = emb.shape[1] == 8 and hid == 128  # fused CUDA-core kernels for the K = 8 layer
        if fast:
            # (no pre-split low part: handing `a_lo` to k_gemm3x collapses its producer pipeline to one
            # piece in flight -- 1.7x slower in-step, VERDICT r01 / profiles/r01_gemm_roles.txt)
            h = torch.empty((E, hid), dtype=emb.dtype, device=emb.device)
            ops.mlp_hidden_fwd(emb, w1s, h, None)
        else:
            h = torch.nn.functional.silu(torch.mm(emb, w1s))
        out = torch.empty((E, W), dtype=emb.dtype, device=emb.device)
        fwd.run(h, out, E)
        ctx.bwd, ctx.w1s, ctx.fast = bwd, w1s, fast
        ctx.save_for_backward(emb)  # the pre-activation is recomputed in the backward (8 FMAs per value)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gw):
        (emb,) = ctx.saved_tensors
        E, hid = emb.shape[0], ctx.w1s.shape[1]
        gh = torch.empty((E, hid), dtype=emb.dtype, device=emb.device)
        ctx.bwd.run(gw.contiguous(), gh, E)
        if ctx.fast:
            gemb = torch.empty_like(emb)
            ops.mlp_hidden_bwd(emb, ctx.w1s, gh, gemb)
            return gemb, None, None, None, None
        pre = torch.mm(emb, ctx.w1s)
        gpre = torch.ops.aten.silu_backward(gh, pre)
        return torch.mm(gpre, ctx.w1s.t()), None, None, None, None


class RadialMLPGemm:
    def __init__(self, lin1, lin2, device):
        self.w1s = (lin1.weight.detach() * lin1.alpha).contiguous()
        hid, W = lin2.weight.shape
        self.W = W
        a2 = float(lin2.alpha)
        self.fwd = ops.GroupedGemm([ops.GemmProblem(0, hid, 0, W, lin2.weight.detach(), scale=a2)], device)
        self.bwd = ops.GroupedGemm([ops.GemmProblem(0, W, 0, hid, lin2.weight.detach(), scale=a2, transposed=True)], device)

    @staticmethod
    def supported(lin1, lin2, dtype) -> bool:
        return dtype == torch.float32 and lin2.weight.shape[0] % 4 == 0 and lin2.weight.shape[1] % 4 == 0

    def __call__(self, emb):
        return _RadialMLPGemmFn.apply(emb.contiguous(), self.w1s, self.fwd, self.bwd, self.W)


# ---------------------------------------------------------------------------------------
class _FusedRadialTPFn(torch.autograd.Function):
    """``out = scatter(TP(x[src], y, silu(emb @ W1 a1) @ W2 a2))`` with the last radial layer fused into the
    tensor-product kernel (forward: the [E, W] weights are produced in tensor memory and consumed in place;
    they are written once on the side only when a backward pass will need them)."""

    @staticmethod
    def forward(ctx, emb, x, y, edge_src, mod: "FusedRadialTP", csr):
        E, hid = emb.shape[0], mod.w1s.shape[1]
        if emb.shape[1] == 8 and hid == 128:
            h = torch.empty((E, hid), dtype=emb.dtype, device=emb.device)
            ops.mlp_hidden_fwd(emb, mod.w1s, h, None)
        else:
            h = torch.nn.functional.silu(torch.mm(emb, mod.w1s))
        need_bwd = any(ctx.needs_input_grad[:3])
        out, w = ops.tp_fused_fwd(mod.fw, x, y, h, edge_src, csr, want_w=need_bwd)
        ctx.mod, ctx.csr = mod, csr
        if need_bwd:
            ctx.save_for_backward(emb, x, y, w, edge_src)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        emb, x, y, w, edge_src = ctx.saved_tensors
        mod = ctx.mod
        gx, gy, gw = ops.tp_scatter_bwd_raw(mod.plan, x, y, w, edge_src, ctx.csr, gout, need_x=ctx.needs_input_grad[1])
        gemb = None
        if ctx.needs_input_grad[0]:
            E, hid = emb.shape[0], mod.w1s.shape[1]
            gh = torch.empty((E, hid), dtype=emb.dtype, device=emb.device)
            mod.bwd.run(gw, gh, E)
            if emb.shape[1] == 8 and hid == 128:
                gemb = torch.empty_like(emb)
                ops.mlp_hidden_bwd(emb, mod.w1s, gh, gemb)
            else:
                pre = torch.mm(emb, mod.w1s)
                gemb = torch.mm(torch.ops.aten.silu_backward(gh, pre), mod.w1s.t())
        return gemb, gx, (gy if ctx.needs_input_grad[2] else None), None, None, None


class FusedRadialTP:
    """Radial MLP (one hidden layer) + TensorProductScatter of one interaction layer as a single forward kernel
    (``nqb_tp_fused_fwd``); backward = ``nqb_tp_scatter_bwd`` + the grouped GEMM for ``grad_h`` + the hidden layer."""

    def __init__(self, lin1, lin2, plan: ops.TPPlan, device):
        self.plan = plan
        self.w1s = (lin1.weight.detach() * lin1.alpha).contiguous()
        hid, W = lin2.weight.shape
        a2 = float(lin2.alpha)
        self.fw = ops.FusedTPWeights(plan, lin2.weight.detach(), a2, device)
        self.bwd = ops.GroupedGemm([ops.GemmProblem(0, W, 0, hid, lin2.weight.detach(), scale=a2, transposed=True)], device)

    @staticmethod
    def supported(lin1, lin2, plan: ops.TPPlan, dtype) -> bool:
        hid, W = lin2.weight.shape
        if dtype != torch.float32 or hid > 128 or hid % 8 or W != plan.weight_numel or W % 4:
            return False
        return int(ops._capi.lib().nqb_tp_fused_slices(plan.handle)) > 0

    def __call__(self, emb, x, y, edge_dst, edge_src):
        csr = ops.csr_cache.get(edge_dst.long().contiguous() if edge_dst.dtype != torch.int64 else edge_dst, x.shape[0])
        if csr.perm is not None:
            return None  # unsorted neighbour list: the caller uses the unfused kernels (which take the permutation)
        return _FusedRadialTPFn.apply(emb.contiguous(), x.contiguous(), y.contiguous(), edge_src.long().contiguous(), self, csr)
