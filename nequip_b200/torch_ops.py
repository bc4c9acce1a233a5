"""``torch.library`` registration of the fused tensor product + scatter (custom ops ``nequip_b200::*``).

What the reference needs from a kernel back-end at the ``TensorProductScatter`` seam
(nequip/nn/_tp_scatter_oeq.py:5-57, nequip/utils/fx.py:52-119,
nequip/nn/compile.py:168-191, nequip/nn/grad_output.py:217-221):

* a traceable op: ``make_fx(tracing_mode="symbolic")`` / ``torch.compile`` / ``torch.export`` see an opaque
  ``torch.ops.nequip_b200.tp_scatter`` node with a fake (meta) implementation -- no ctypes calls inside the trace;
* autograd of any order: forces are a first derivative, a force loss needs the derivative of that
  (``create_graph=self.training``).  The product is trilinear in (x, Y, w), so every derivative of the backward
  op is again a sum of forward / backward kernel calls with one argument replaced by a cotangent:
  with  L = <ggx, gx> + <ggy, gy> + <ggw, gw>  and  (gx, gy, gw) = B(g; x, y, w):
      dL/dg = T(ggx, y, w) + T(x, ggy, w) + T(x, y, ggw)
      dL/dx = B_x(g; x, ggy, w) + B_x(g; x, y, ggw),   dL/dy, dL/dw analogously.
  Both ops below are registered with ``register_autograd`` in terms of each other, so the chain never ends.

Plans (immutable kernel bindings) are looked up by a string key, because custom-op arguments must be
tensors / scalars / strings.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops

_PLANS: Dict[str, "ops.TPPlan"] = {}


def register_plan(plan: "ops.TPPlan") -> str:
    key = plan.sig.canonical() + "|" + plan.opts.tag()
    _PLANS[key] = plan
    return key


def _plan(key: str) -> "ops.TPPlan":
    try:
        return _PLANS[key]
    except KeyError:
        raise RuntimeError(f"nequip_b200::tp_scatter: unknown plan key {key!r} (register_plan() first)") from None


def _csr(edge_dst: torch.Tensor, n: int) -> "ops.EdgeCSR":
    return ops.csr_cache.get(edge_dst if edge_dst.dtype == torch.int64 else edge_dst.long(), n)


@torch.library.custom_op("nequip_b200::tp_scatter", mutates_args=())
def tp_scatter(x: torch.Tensor, edge_attr: torch.Tensor, edge_weight: torch.Tensor, edge_dst: torch.Tensor,
               edge_src: torch.Tensor, plan_key: str) -> torch.Tensor:
    plan = _plan(plan_key)
    ops._require_cuda(x, edge_attr, edge_weight, edge_dst, edge_src)
    if x.dtype not in ops._DT:
        raise TypeError(f"nequip_b200::tp_scatter: unsupported dtype {x.dtype}")
    E = edge_src.numel()
    if x.dim() != 2 or x.shape[1] != plan.d_in or tuple(edge_attr.shape) != (E, plan.s_dim) or \
            tuple(edge_weight.shape) != (E, plan.weight_numel) or edge_dst.numel() != E:
        raise ValueError(f"nequip_b200::tp_scatter: expected x [N,{plan.d_in}], edge_attr [{E},{plan.s_dim}], "
                         f"edge_weight [{E},{plan.weight_numel}]; got {tuple(x.shape)}, {tuple(edge_attr.shape)}, "
                         f"{tuple(edge_weight.shape)}")
    dt = x.dtype
    x = x.contiguous()
    y = edge_attr.to(dt).contiguous()
    w = edge_weight.to(dt).contiguous()
    src = edge_src.long().contiguous()
    dst = edge_dst.long().contiguous()
    N = x.shape[0]
    out = torch.empty((N, plan.d_out), dtype=dt, device=x.device)
    if N > 0:
        csr = _csr(dst, N)
        ops._capi.check(
            ops._capi.lib().nqb_tp_scatter_fwd(plan.handle, ops._DT[dt], ops._ptr(x), ops._ptr(y), ops._ptr(w),
                                               ops._ptr(csr.row_ptr), ops._ptr(csr.perm), ops._ptr(src), N, E,
                                               ops._ptr(out), ops._stream()), "nqb_tp_scatter_fwd")
    return out


@tp_scatter.register_fake
def _(x, edge_attr, edge_weight, edge_dst, edge_src, plan_key):
    return x.new_empty((x.shape[0], _plan(plan_key).d_out))


@torch.library.custom_op("nequip_b200::tp_scatter_bwd", mutates_args=())
def tp_scatter_bwd(grad_out: torch.Tensor, x: torch.Tensor, edge_attr: torch.Tensor, edge_weight: torch.Tensor,
                   edge_dst: torch.Tensor, edge_src: torch.Tensor, plan_key: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    plan = _plan(plan_key)
    dt = x.dtype
    x = x.contiguous()
    y = edge_attr.to(dt).contiguous()
    w = edge_weight.to(dt).contiguous()
    src = edge_src.long().contiguous()
    dst = edge_dst.long().contiguous()
    N = x.shape[0]
    csr = _csr(dst, N) if N > 0 else None
    gx, gy, gw = ops.tp_scatter_bwd_raw(plan, x, y, w, src, csr, grad_out.to(dt), need_x=True)
    return gx, gy, gw


@tp_scatter_bwd.register_fake
def _(grad_out, x, edge_attr, edge_weight, edge_dst, edge_src, plan_key):
    return torch.empty_like(x), torch.empty_like(edge_attr), torch.empty_like(edge_weight)


def _fwd_setup(ctx, inputs, output):
    x, y, w, dst, src, key = inputs
    ctx.key = key
    ctx.save_for_backward(x, y, w, dst, src)


def _fwd_backward(ctx, gout):
    x, y, w, dst, src = ctx.saved_tensors
    gx, gy, gw = torch.ops.nequip_b200.tp_scatter_bwd(gout, x, y, w, dst, src, ctx.key)
    return gx, gy, gw, None, None, None


tp_scatter.register_autograd(_fwd_backward, setup_context=_fwd_setup)


def _bwd_setup(ctx, inputs, output):
    g, x, y, w, dst, src, key = inputs
    ctx.key = key
    ctx.save_for_backward(g, x, y, w, dst, src)


def _bwd_backward(ctx, ggx, ggy, ggw):
    g, x, y, w, dst, src = ctx.saved_tensors
    key = ctx.key
    T = torch.ops.nequip_b200.tp_scatter
    B = torch.ops.nequip_b200.tp_scatter_bwd
    d_g = d_x = d_y = d_w = None

    def acc(a, b):
        return b if a is None else a + b

    if ggx is not None:
        d_g = acc(d_g, T(ggx, y, w, dst, src, key))
        _, ay, aw = B(g, ggx, y, w, dst, src, key)
        d_y, d_w = acc(d_y, ay), acc(d_w, aw)
    if ggy is not None:
        d_g = acc(d_g, T(x, ggy, w, dst, src, key))
        bx, _, bw = B(g, x, ggy, w, dst, src, key)
        d_x, d_w = acc(d_x, bx), acc(d_w, bw)
    if ggw is not None:
        d_g = acc(d_g, T(x, y, ggw, dst, src, key))
        cx, cy, _ = B(g, x, y, ggw, dst, src, key)
        d_x, d_y = acc(d_x, cx), acc(d_y, cy)
    return d_g, d_x, d_y, d_w, None, None, None


tp_scatter_bwd.register_autograd(_bwd_backward, setup_context=_bwd_setup)

__all__ = ["tp_scatter", "tp_scatter_bwd", "register_plan"]
