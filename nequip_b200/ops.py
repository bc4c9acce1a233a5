"""Torch-facing operators over the C ABI (device memory, streams and autograd glue only).

``tp_scatter``  -- the fused tensor product + scatter that replaces
                   ``TensorProductScatter.forward`` (nequip/nn/_tp_scatter_base.py:35-38)
                   and its autograd (first order: what forces need,
                   nequip/nn/grad_output.py:217-221 with ``create_graph=False``).
``spherical_harmonics`` / ``edge_embed`` -- the edge-embedding kernels replacing
                   nequip/nn/embedding/_edge.py:65-80,136-150,193-198 + nequip/nn/utils.py:68-118.

There is no CPU or eager fallback: CPU tensors raise, a missing/unbuildable kernel
library raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _capi, build
from .codegen import GenOptions, TPSignature
from .irreps import Irreps

_DT = {torch.float32: 0, torch.float64: 1}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _require_cuda(*ts: torch.Tensor):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "nequip_b200 kernels run on CUDA (sm_100a) only; got a CPU tensor and there is no CPU fallback"
            )


# ---------------------------------------------------------------------------------------
# plans
# ---------------------------------------------------------------------------------------
class TPPlan:
    """Immutable binding of one TensorProductScatter signature to its kernel library."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, opts: Optional[GenOptions] = None):
        self.sig = TPSignature(Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out), list(instructions))
        self.opts = opts or GenOptions()
        self.spec_path = build.ensure_spec(self.sig, self.opts)
        L = _capi.lib()

        def arr(irr):
            a = (_capi.NqbIrrep * len(irr))()
            for i, (mul, ir) in enumerate(irr):
                a[i].mul, a[i].l, a[i].p = mul, ir.l, ir.p
            return a

        in1, in2, out = arr(self.sig.irreps_in1), arr(self.sig.irreps_in2), arr(self.sig.irreps_out)
        ins = (_capi.NqbInstruction * len(self.sig.instructions))()
        for i, (a, b, c) in enumerate(self.sig.instructions):
            ins[i].i_in1, ins[i].i_in2, ins[i].i_out = a, b, c
        h = C.c_void_p()
        _capi.check(
            L.nqb_plan_create(in1, len(in1), in2, len(in2), out, len(out), ins, len(ins),
                              self.spec_path.encode(), C.byref(h)),
            "nqb_plan_create",
        )
        self._h = h
        self.d_in, self.s_dim = self.sig.d_in, self.sig.s_dim
        self.weight_numel, self.d_out = self.sig.weight_numel, self.sig.d_out

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _capi.lib().nqb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


_plans: Dict[str, TPPlan] = {}
_plans_lock = threading.Lock()


def get_plan(irreps_in1, irreps_in2, irreps_out, instructions, opts: Optional[GenOptions] = None) -> TPPlan:
    opts = opts or GenOptions()
    sig = TPSignature(Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out), list(instructions))
    key = sig.canonical() + "|" + opts.tag()
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = TPPlan(irreps_in1, irreps_in2, irreps_out, instructions, opts)
            _plans[key] = p
        return p


# ---------------------------------------------------------------------------------------
# destination CSR
# ---------------------------------------------------------------------------------------
@dataclass
class EdgeCSR:
    row_ptr: torch.Tensor  # [N+1] int64
    perm: Optional[torch.Tensor]  # [E] int64 (slot -> edge id) or None when edges are already grouped
    num_nodes: int
    num_edges: int


def build_csr(edge_dst: torch.Tensor, num_nodes: int, assume_sorted: Optional[bool] = None) -> EdgeCSR:
    """CSR over edges grouped by destination.  ``assume_sorted=None`` checks on the
    device (one host sync); the reference's neighbour lists are grouped by centre atom
    (nequip/data/transforms/neighborlist.py:120-157) so the common case needs no sort."""
    _require_cuda(edge_dst)
    if edge_dst.dtype != torch.int64:
        edge_dst = edge_dst.long()
    edge_dst = edge_dst.contiguous()
    E = edge_dst.numel()
    L = _capi.lib()
    st = _stream()
    is_sorted = assume_sorted
    deferred = getattr(_sorted_tls, "flags", None)
    if is_sorted is None and deferred is not None:
        # CUDA-graph capture (nequip_b200/graph.py): no host sync allowed -- run the check kernel, keep its
        # flag for the caller to verify after the replay, and build the CSR as if sorted
        flag = torch.empty(1, dtype=torch.int32, device=edge_dst.device)
        _capi.check(L.nqb_csr_check_sorted(_ptr(edge_dst), E, _ptr(flag), st), "nqb_csr_check_sorted")
        deferred.append(flag)
        is_sorted = True
    if is_sorted is None:
        flag = torch.empty(1, dtype=torch.int32, device=edge_dst.device)
        _capi.check(L.nqb_csr_check_sorted(_ptr(edge_dst), E, _ptr(flag), st), "nqb_csr_check_sorted")
        is_sorted = bool(flag.item())
    perm = None
    keys = edge_dst
    if not is_sorted:
        keys, perm = torch.sort(edge_dst, stable=True)
        perm = perm.contiguous()
        keys = keys.contiguous()
    row_ptr = torch.empty(num_nodes + 1, dtype=torch.int64, device=edge_dst.device)
    _capi.check(L.nqb_csr_from_sorted(_ptr(keys), E, num_nodes, _ptr(row_ptr), st), "nqb_csr_from_sorted")
    return EdgeCSR(row_ptr, perm, num_nodes, E)


_sorted_tls = threading.local()


class deferred_sorted_check:
    """Context manager: inside it ``build_csr`` does not synchronise to learn whether the edge list is grouped
    by destination; it records the device flags (1 = sorted) in ``self.flags`` and assumes sorted."""

    def __enter__(self):
        self.flags: List[torch.Tensor] = []
        self._prev = getattr(_sorted_tls, "flags", None)
        _sorted_tls.flags = self.flags
        return self

    def __exit__(self, *exc):
        _sorted_tls.flags = self._prev
        return False


class _CSRCache:
    """Per-thread one-entry cache: all layers of one forward share the same edge_index."""

    def __init__(self):
        self._tls = threading.local()

    def get(self, edge_dst: torch.Tensor, num_nodes: int) -> EdgeCSR:
        key = (edge_dst.data_ptr(), edge_dst.numel(), edge_dst._version, num_nodes, edge_dst.device)
        ent = getattr(self._tls, "ent", None)
        if ent is not None and ent[0] == key:
            return ent[1]
        csr = build_csr(edge_dst, num_nodes)
        # hold a reference to edge_dst so the data_ptr cannot be recycled while cached
        self._tls.ent = (key, csr, edge_dst)
        return csr

    def clear(self):
        self._tls.ent = None


csr_cache = _CSRCache()


# ---------------------------------------------------------------------------------------
# fused TP + scatter
# ---------------------------------------------------------------------------------------
class _TPScatterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, edge_attr, edge_weight, edge_src, plan: TPPlan, csr: EdgeCSR):
        L = _capi.lib()
        N, E = x.shape[0], edge_src.numel()
        out = torch.empty((N, plan.d_out), dtype=x.dtype, device=x.device)
        if N > 0:
            _capi.check(
                L.nqb_tp_scatter_fwd(plan.handle, _DT[x.dtype], _ptr(x), _ptr(edge_attr), _ptr(edge_weight),
                                     _ptr(csr.row_ptr), _ptr(csr.perm), _ptr(edge_src), N, E, _ptr(out), _stream()),
                "nqb_tp_scatter_fwd",
            )
        ctx.plan, ctx.csr = plan, csr
        ctx.save_for_backward(x, edge_attr, edge_weight, edge_src)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        x, y, w, edge_src = ctx.saved_tensors
        gx, gy, gw = tp_scatter_bwd_raw(ctx.plan, x, y, w, edge_src, ctx.csr, gout, need_x=ctx.needs_input_grad[0])
        return gx, (gy if ctx.needs_input_grad[1] else None), (gw if ctx.needs_input_grad[2] else None), None, None, None


def tp_scatter(plan: TPPlan, x, edge_attr, edge_weight, edge_dst, edge_src, csr: Optional[EdgeCSR] = None):
    """``out[n] = sum_{e: dst[e]=n} TP_uvu(x[src[e]], edge_attr[e], edge_weight[e])`` -> ``[x.size(0), D_mid]``."""
    _require_cuda(x, edge_attr, edge_weight, edge_dst, edge_src)
    if x.dtype not in _DT:
        raise TypeError(f"nequip_b200.tp_scatter: unsupported dtype {x.dtype}")
    dt = x.dtype
    if x.dim() != 2 or x.shape[1] != plan.d_in:
        raise ValueError(f"x must be [N, {plan.d_in}], got {tuple(x.shape)}")
    E = edge_src.numel()
    if tuple(edge_attr.shape) != (E, plan.s_dim) or tuple(edge_weight.shape) != (E, plan.weight_numel):
        raise ValueError(
            f"edge_attr/edge_weight must be [{E}, {plan.s_dim}] / [{E}, {plan.weight_numel}], got "
            f"{tuple(edge_attr.shape)} / {tuple(edge_weight.shape)}"
        )
    if edge_dst.numel() != E:
        raise ValueError("edge_dst and edge_src differ in length")
    x = x.contiguous()
    y = edge_attr.to(dt).contiguous()
    w = edge_weight.to(dt).contiguous()
    src = edge_src.long().contiguous()
    if csr is None:
        csr = csr_cache.get(edge_dst.long().contiguous() if edge_dst.dtype != torch.int64 else edge_dst, x.shape[0])
    elif csr.num_nodes != x.shape[0] or csr.num_edges != E:
        raise ValueError("EdgeCSR does not match x / edge_index")
    return _TPScatterFn.apply(x, y, w, src, plan, csr)


# ---------------------------------------------------------------------------------------
# fused last radial-MLP layer + TP + scatter (forward) -- nqb_tp_fused_fwd
# ---------------------------------------------------------------------------------------
class FusedTPWeights:
    """Second-layer radial-MLP weights ``W2 [K, W]`` (times ``alpha2``) permuted into the slice order of the
    signature's fused kernel, split hi/lo and laid out for tensor memory (once per model), plus the split of
    the grid (one CTA per SM) over the slices in proportion to their cost."""

    def __init__(self, plan: TPPlan, W2: torch.Tensor, alpha2: float, device):
        from .codegen import TPGenerator

        L = _capi.lib()
        self.plan = plan
        lay = TPGenerator(plan.sig, plan.opts).fused_layout()
        nslice = int(L.nqb_tp_fused_slices(plan.handle))
        if lay is None or nslice == 0 or nslice != len(lay["slices"]):
            raise RuntimeError("FusedTPWeights: this signature has no fused kernel")
        K, W = int(W2.shape[0]), int(W2.shape[1])
        if W != plan.weight_numel or K > 128 or K % 8:
            raise ValueError(f"FusedTPWeights: W2 must be [K <= 128 (multiple of 8), {plan.weight_numel}], got {tuple(W2.shape)}")
        self.K, self.nslice = K, nslice
        cols = torch.tensor(lay["cols"], dtype=torch.long, device=device)
        W2d = W2.detach().to(device=device, dtype=torch.float32)
        Wp = torch.zeros((K, cols.numel()), dtype=torch.float32, device=device)
        ok = cols >= 0
        Wp[:, ok] = W2d[:, cols[ok]]
        self.prepared = torch.empty(int(L.nqb_gemm_t_prepared_floats(K, Wp.shape[1])), dtype=torch.float32, device=device)
        _capi.check(L.nqb_gemm_t_prepare(_ptr(Wp), Wp.shape[1], K, Wp.shape[1], 0, float(alpha2), _ptr(self.prepared),
                                         _stream()), "nqb_gemm_t_prepare")
        G = torch.cuda.get_device_properties(device).multi_processor_count
        self.cta0, self.nctas = self.split_grid(lay["cost"], G)
        self.cta0_dev = torch.tensor(self.cta0, dtype=torch.int32, device=device)

    @staticmethod
    def split_grid(cost, G: int):
        """CTAs per slice proportional to cost (every slice gets at least one); returns (prefix, total)."""
        S = len(cost)
        n = [1] * S
        for _ in range(max(0, G - S)):
            i = max(range(S), key=lambda j: cost[j] / n[j])
            n[i] += 1
        pre = [0]
        for v in n:
            pre.append(pre[-1] + v)
        return pre, pre[-1]


def tp_fused_fwd(fw: FusedTPWeights, x: torch.Tensor, y: torch.Tensor, h: torch.Tensor, edge_src: torch.Tensor,
                 csr: EdgeCSR, want_w: bool):
    """``out [N, D_mid]`` (and the per-edge weights ``w [E, W]`` when ``want_w``) of the fused kernel."""
    _require_cuda(x, y, h, edge_src)
    plan = fw.plan
    if csr.perm is not None:
        raise RuntimeError("tp_fused_fwd: edges must be grouped by destination (no permutation)")
    if x.dtype != torch.float32 or y.dtype != torch.float32 or h.dtype != torch.float32:
        raise TypeError("tp_fused_fwd: float32 only")
    N, E = x.shape[0], edge_src.numel()
    if x.shape[1] != plan.d_in or tuple(y.shape) != (E, plan.s_dim) or h.shape[0] != E or h.shape[1] != fw.K:
        raise ValueError("tp_fused_fwd: shape mismatch")
    x, y, h = x.contiguous(), y.contiguous(), h.contiguous()
    out = torch.empty((N, plan.d_out), dtype=torch.float32, device=x.device)
    w = torch.empty((E, plan.weight_numel), dtype=torch.float32, device=x.device) if want_w else None
    _capi.check(
        _capi.lib().nqb_tp_fused_fwd(plan.handle, _ptr(x), _ptr(y), _ptr(h), h.stride(0), fw.K, _ptr(fw.prepared),
                                     _ptr(csr.row_ptr), _ptr(edge_src), N, E, _ptr(out), _ptr(w), _ptr(fw.cta0_dev),
                                     int(fw.nctas), _stream()),
        "nqb_tp_fused_fwd",
    )
    return out, w


_DETERMINISTIC = os.environ.get("NQB_DETERMINISTIC", "0") not in ("", "0")


def set_deterministic(on: bool = True) -> None:
    """Bitwise-repeatable backward of the fused TP+scatter (also implied by ``torch.use_deterministic_algorithms``):
    grad_x goes through a per-edge buffer and a source-sorted segmented sum, grad_Y through one slice per writer,
    instead of ``red.global.add`` in arrival order (the default, like the reference's OpenEquivariance back-end)."""
    global _DETERMINISTIC
    _DETERMINISTIC = bool(on)


def deterministic() -> bool:
    return _DETERMINISTIC or torch.are_deterministic_algorithms_enabled()


class _SrcCSRCache:
    """Source-sorted view of the edge list (the reference's edge_transpose_perm): one entry per thread."""

    def __init__(self):
        self._tls = threading.local()

    def get(self, edge_src: torch.Tensor, num_nodes: int):
        key = (edge_src.data_ptr(), edge_src.numel(), edge_src._version, num_nodes, edge_src.device)
        ent = getattr(self._tls, "ent", None)
        if ent is not None and ent[0] == key:
            return ent[1], ent[2]
        keys, perm = torch.sort(edge_src, stable=True)
        seg = torch.empty(num_nodes + 1, dtype=torch.int64, device=edge_src.device)
        _capi.check(_capi.lib().nqb_csr_from_sorted(_ptr(keys.contiguous()), edge_src.numel(), num_nodes, _ptr(seg), _stream()),
                    "nqb_csr_from_sorted")
        self._tls.ent = (key, perm.contiguous(), seg, edge_src)
        return self._tls.ent[1], seg

    def clear(self):
        self._tls.ent = None


src_csr_cache = _SrcCSRCache()


def tp_scatter_bwd_raw(plan: TPPlan, x, y, w, edge_src, csr: EdgeCSR, gout, need_x: bool = True,
                       force_deterministic: Optional[bool] = None):
    """Backward kernels of the fused TP+scatter on raw tensors: (grad_x or None, grad_y, grad_w)."""
    L = _capi.lib()
    N, E = x.shape[0], edge_src.numel()
    det = deterministic() if force_deterministic is None else bool(force_deterministic)
    gw = torch.empty_like(w)
    if N == 0 or E == 0:
        gw.zero_()
        return (torch.zeros_like(x) if need_x else None), torch.zeros_like(y), gw
    gout = gout.contiguous()
    dt = _DT[x.dtype]
    if not det:
        gx = torch.zeros_like(x) if need_x else None
        gy = torch.zeros_like(y)
        _capi.check(L.nqb_tp_scatter_bwd(plan.handle, dt, _ptr(x), _ptr(y), _ptr(w), _ptr(csr.row_ptr), _ptr(csr.perm),
                                         _ptr(edge_src), _ptr(gout), N, E, _ptr(gx), _ptr(gy), _ptr(gw), 0, _stream()),
                    "nqb_tp_scatter_bwd")
        return gx, gy, gw
    ns = int(L.nqb_tp_scatter_gy_slices(plan.handle, dt))
    if ns <= 0:
        raise RuntimeError("nequip_b200: this kernel library has no deterministic backward")
    gxe = torch.empty((E, x.shape[1]), dtype=x.dtype, device=x.device) if need_x else None
    gys = torch.zeros((ns, E, y.shape[1]), dtype=y.dtype, device=y.device)
    _capi.check(L.nqb_tp_scatter_bwd(plan.handle, dt, _ptr(x), _ptr(y), _ptr(w), _ptr(csr.row_ptr), _ptr(csr.perm),
                                     _ptr(edge_src), _ptr(gout), N, E, _ptr(gxe), _ptr(gys), _ptr(gw), 1, _stream()),
                "nqb_tp_scatter_bwd")
    gx = None
    if need_x:
        perm_t, seg_t = src_csr_cache.get(edge_src, N)
        gx = torch.empty_like(x)
        _capi.check(L.nqb_segment_sum(dt, _ptr(gxe), x.shape[1], _ptr(perm_t), _ptr(seg_t), N, _ptr(gx), _stream()),
                    "nqb_segment_sum")
    gy = gys[0]
    for q in range(1, ns):  # fixed order
        gy = gy + gys[q]
    return gx, gy, gw


# ---------------------------------------------------------------------------------------
# spherical harmonics / edge embedding
# ---------------------------------------------------------------------------------------
class _SHFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, lmax: int, out_dtype):
        L = _capi.lib()
        E = vec.shape[0]
        y = torch.empty((E, (lmax + 1) ** 2), dtype=out_dtype, device=vec.device)
        _capi.check(L.nqb_sh_fwd(lmax, _ptr(vec), E, _DT[out_dtype], _ptr(y), _stream()), "nqb_sh_fwd")
        ctx.lmax, ctx.out_dtype = lmax, out_dtype
        ctx.save_for_backward(vec)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        (vec,) = ctx.saved_tensors
        L = _capi.lib()
        E = vec.shape[0]
        gvec = torch.empty_like(vec)
        gy = gy.to(ctx.out_dtype).contiguous()
        _capi.check(L.nqb_sh_bwd(ctx.lmax, _ptr(vec), E, _DT[ctx.out_dtype], _ptr(gy), _ptr(gvec), _stream()),
                    "nqb_sh_bwd")
        return gvec, None, None


def spherical_harmonics(vec: torch.Tensor, lmax: int, out_dtype=torch.float32) -> torch.Tensor:
    """``o3.SphericalHarmonics(lmax, normalize=True, "component")`` of ``[E,3]`` float64 edge vectors."""
    _require_cuda(vec)
    if vec.dim() != 2 or vec.shape[1] != 3:
        raise ValueError("vec must be [E, 3]")
    return _SHFn.apply(vec.double().contiguous(), int(lmax), out_dtype)


class _EdgeEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, edge_index, shift, cell, lmax, num_bessel, r_max, poly_p, prefactor, out_dtype, sink=None):
        L = _capi.lib()
        N, E = pos.shape[0], edge_index.shape[1]
        dev = pos.device
        vec = torch.empty((E, 3), dtype=torch.float64, device=dev)
        y = torch.empty((E, (lmax + 1) ** 2), dtype=out_dtype, device=dev)
        emb = torch.empty((E, num_bessel), dtype=out_dtype, device=dev)
        _capi.check(
            L.nqb_edge_embed_fwd(lmax, num_bessel, r_max, poly_p, prefactor, _ptr(pos), _ptr(edge_index), _ptr(shift),
                                 _ptr(cell), N, E, _DT[out_dtype], _ptr(vec), _ptr(y), _ptr(emb), _stream()),
            "nqb_edge_embed_fwd",
        )
        ctx.args = (lmax, num_bessel, r_max, poly_p, prefactor, out_dtype, N)
        ctx.sink = sink
        ctx.save_for_backward(vec, edge_index)
        ctx.mark_non_differentiable(vec)
        return vec, y, emb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _gvec, gy, gemb):
        vec, edge_index = ctx.saved_tensors
        lmax, num_bessel, r_max, poly_p, prefactor, out_dtype, N = ctx.args
        L = _capi.lib()
        E = vec.shape[0]
        gpos = torch.zeros((N, 3), dtype=torch.float64, device=vec.device)
        gy = None if gy is None else gy.to(out_dtype).contiguous()
        gemb = None if gemb is None else gemb.to(out_dtype).contiguous()
        # the per-edge gradient dE/d(edge vector) is what the virial is made of (sum_e r_e (x) g_e): keep it
        # when the caller asked for it
        gvec = torch.empty_like(vec) if ctx.sink is not None else None
        _capi.check(
            L.nqb_edge_embed_bwd(lmax, num_bessel, r_max, poly_p, prefactor, _ptr(vec), _ptr(edge_index), N, E,
                                 _DT[out_dtype], _ptr(gy), _ptr(gemb), _ptr(gpos), _ptr(gvec), _stream()),
            "nqb_edge_embed_bwd",
        )
        if ctx.sink is not None:
            ctx.sink["edge_vectors"], ctx.sink["edge_vector_grad"] = vec, gvec
        return gpos, None, None, None, None, None, None, None, None, None, None


class _EdgeEmbedVecFn(torch.autograd.Function):
    """Harmonics + radial embedding of GIVEN edge vectors (the ML-IAP branch: LAMMPS hands over the vectors,
    nequip/nn/grad_output.py:270-296); backward = dE/d(edge vectors)."""

    @staticmethod
    def forward(ctx, vec, lmax, num_bessel, r_max, poly_p, prefactor, out_dtype):
        L = _capi.lib()
        E = vec.shape[0]
        dev = vec.device
        # the fused kernel computes pos[src] - pos[dst]: atoms 0..E-1 are the vectors, atom E is the origin
        pos = torch.cat([vec, torch.zeros((1, 3), dtype=torch.float64, device=dev)], 0)
        ar = torch.arange(E, dtype=torch.int64, device=dev)
        edge_index = torch.stack([torch.full_like(ar, E), ar]).contiguous()
        vec_out = torch.empty((E, 3), dtype=torch.float64, device=dev)
        y = torch.empty((E, (lmax + 1) ** 2), dtype=out_dtype, device=dev)
        emb = torch.empty((E, num_bessel), dtype=out_dtype, device=dev)
        _capi.check(
            L.nqb_edge_embed_fwd(lmax, num_bessel, r_max, poly_p, prefactor, _ptr(pos), _ptr(edge_index), 0, 0, E + 1, E,
                                 _DT[out_dtype], _ptr(vec_out), _ptr(y), _ptr(emb), _stream()),
            "nqb_edge_embed_fwd",
        )
        ctx.args = (lmax, num_bessel, r_max, poly_p, prefactor, out_dtype)
        ctx.save_for_backward(vec_out, edge_index)
        return y, emb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, gemb):
        vec, edge_index = ctx.saved_tensors
        lmax, num_bessel, r_max, poly_p, prefactor, out_dtype = ctx.args
        E = vec.shape[0]
        gvec = torch.empty_like(vec)
        gy = None if gy is None else gy.to(out_dtype).contiguous()
        gemb = None if gemb is None else gemb.to(out_dtype).contiguous()
        _capi.check(
            _capi.lib().nqb_edge_embed_bwd(lmax, num_bessel, r_max, poly_p, prefactor, _ptr(vec), _ptr(edge_index), E + 1, E,
                                           _DT[out_dtype], _ptr(gy), _ptr(gemb), 0, _ptr(gvec), _stream()),
            "nqb_edge_embed_bwd",
        )
        return gvec, None, None, None, None, None, None


def edge_embed_from_vectors(vec, *, lmax: int, num_bessel: int = 8, r_max: float, poly_p: float = 6.0,
                            prefactor: float = 1.0, out_dtype=torch.float32):
    """``(edge_attrs [E,(lmax+1)^2], edge_embedding [E,num_bessel])`` of given ``[E,3]`` edge vectors,
    differentiable w.r.t. the vectors."""
    _require_cuda(vec)
    return _EdgeEmbedVecFn.apply(vec.double().contiguous(), int(lmax), int(num_bessel), float(r_max), float(poly_p),
                                 float(prefactor), out_dtype)


def edge_embed(pos, edge_index, shift=None, cell=None, *, lmax: int, num_bessel: int = 8, r_max: float,
               poly_p: float = 6.0, prefactor: float = 1.0, out_dtype=torch.float32, edge_grad_sink=None):
    """Edge vectors, harmonics and Bessel x cutoff embedding in one kernel.

    Returns ``(edge_vectors [E,3] f64, edge_attrs [E,(lmax+1)^2], edge_embedding [E,num_bessel])``.
    Differentiable w.r.t. ``pos`` (forces).  ``edge_grad_sink`` (a dict): the backward pass also stores the
    edge vectors and dE/d(edge vector) in it, from which the virial / cell gradient follows."""
    _require_cuda(pos, edge_index)
    pos = pos.double().contiguous()
    edge_index = edge_index.long().contiguous()
    if (shift is None) != (cell is None):
        raise ValueError("shift and cell must be given together")
    if shift is not None:
        shift = shift.double().contiguous()
        cell = cell.double().reshape(3, 3).contiguous()
    return _EdgeEmbedFn.apply(pos, edge_index, shift, cell, int(lmax), int(num_bessel), float(r_max),
                              float(poly_p), float(prefactor), out_dtype, edge_grad_sink)


# ---------------------------------------------------------------------------------------
# grouped fp32-accurate GEMM on the tensor cores (tcgen05 3xTF32) -- nqb_gemm_grouped
# ---------------------------------------------------------------------------------------
@dataclass
class GemmProblem:
    """C[M, N] (ldc, at c_off) (+)= rowscale[m] * A[M, K] (lda, at a_off) @ B[K, N]."""

    a_off: int
    lda: int
    c_off: int
    ldc: int
    B: torch.Tensor  # [K, N] (or [N, K] when transposed=True), float32
    scale: float = 1.0
    transposed: bool = False
    accumulate: bool = False  # C += ... (read-modify-write; at most ONE problem of the launch may touch an element)
    rs_off: int = -1  # row of the [R, M] row-scale matrix, -1 = none
    skip_zero_rows: bool = False  # rows with row scale 0 are left untouched (disjoint row-masked writers)
    atomic: bool = False  # C += ... with red.global.add (several problems of one launch add into the same C)


class GroupedGemm:
    """A fixed list of GEMM problems sharing M, with their weights prepared once (split hi/lo, tiled)."""

    def __init__(self, problems, device):
        L = _capi.lib()
        self.problems = list(problems)
        if not self.problems:
            raise ValueError("GroupedGemm: empty problem list")
        rows, blobs, b_off, tile0 = [], [], 0, 0
        for p in self.problems:
            K, N = (p.B.shape[1], p.B.shape[0]) if p.transposed else (p.B.shape[0], p.B.shape[1])
            if any(v % 4 for v in (K, N, p.lda, p.ldc, p.a_off, p.c_off)):
                raise ValueError("GroupedGemm: K, N, lda, ldc and offsets must be multiples of 4")
            if p.B.dtype != torch.float32:
                raise TypeError("GroupedGemm: float32 only")
            Bc = p.B.detach().to(device).contiguous()
            nfl = int(L.nqb_gemm_prepared_floats(K, N))
            prep = torch.empty(nfl, dtype=torch.float32, device=device)
            _capi.check(L.nqb_gemm_prepare(_ptr(Bc), Bc.shape[1], K, N, int(p.transposed), float(p.scale), _ptr(prep),
                                           _stream()), "nqb_gemm_prepare")
            kchunks, ntiles = (K + 31) // 32, (N + 127) // 128
            rows.append([p.a_off, p.c_off, b_off, p.rs_off, p.lda, p.ldc, K, N, kchunks, ntiles, tile0,
                         (1 if p.accumulate else 0) | (2 if p.skip_zero_rows else 0) | (4 if p.atomic else 0)])
            blobs.append(prep)
            b_off += nfl
            tile0 += ntiles
        self.ntiles_total = tile0
        self.tile_ctas, self.sched_ctas = self._weighted_split(rows, device)
        self.prepared = torch.cat(blobs)
        self.descs = torch.tensor(rows, dtype=torch.int64, device=device)
        self.ndesc = len(rows)

    @staticmethod
    def _weighted_split(rows, device):
        """CTAs per N-tile proportional to the tile's cost (pieces of K to multiply + columns to store, reduce-adds
        twice a store), for a grid of one CTA per SM.  None when there are more N-tiles than SMs."""
        if torch.device(device).type != "cuda":
            return None, 0
        G = torch.cuda.get_device_properties(device).multi_processor_count
        cost = []
        for (_a, _c, _b, _rs, _lda, _ldc, K, N, _kch, ntiles, _t0, flags) in rows:
            for t in range(ntiles):
                ncols = min(128, N - t * 128)
                cost.append(((K + 31) // 32) * 1000.0 + ((ncols + 31) // 32) * 650.0 * (2.0 if flags & 5 else 1.0))
        T = len(cost)
        # uniform launches (the radial-MLP GEMMs) keep the even split: its CTAs sweep the M-tiles in lockstep
        # and share every A tile in L2
        if T > G or T < 2 or max(cost) < 1.3 * min(cost):
            return None, 0
        n = [1] * T
        for _ in range(G - T):  # give the next CTA to the tile with the largest per-CTA load
            i = max(range(T), key=lambda j: cost[j] / n[j])
            n[i] += 1
        tab, c0 = [], 0
        for j in range(T):
            tab += [c0, n[j]]
            c0 += n[j]
        return torch.tensor(tab, dtype=torch.int32, device=device), c0

    def run(self, a: torch.Tensor, c: torch.Tensor, M: int, rowscale: Optional[torch.Tensor] = None,
            a_lo: Optional[torch.Tensor] = None):
        """``a_lo``: optional pre-split low parts of ``a`` (same shape/strides; see ``mlp_hidden_fwd``)."""
        _require_cuda(a, c)
        if a.dtype != torch.float32 or c.dtype != torch.float32:
            raise TypeError("GroupedGemm.run: float32 only")
        if a_lo is not None and (a_lo.dtype != torch.float32 or a_lo.shape != a.shape or a_lo.stride() != a.stride()):
            raise ValueError("GroupedGemm.run: a_lo must match a")
        _capi.check(
            _capi.lib().nqb_gemm_grouped(_ptr(self.descs), self.ndesc, self.ntiles_total, _ptr(self.tile_ctas),
                                         int(self.sched_ctas), _ptr(a), _ptr(a_lo),
                                         _ptr(self.prepared), _ptr(c), _ptr(rowscale),
                                         (int(rowscale.shape[-1]) if rowscale is not None else 0), int(M), _stream()),
            "nqb_gemm_grouped",
        )
        return c


def mlp_hidden_fwd(emb: torch.Tensor, w1s: torch.Tensor, h: torch.Tensor, h_lo: Optional[torch.Tensor] = None) -> None:
    """``h = silu(emb @ w1s)`` ([E,8] x [8,128]); ``h_lo`` (optional) receives the tf32 low part of ``h``."""
    _require_cuda(emb, w1s, h)
    _capi.check(_capi.lib().nqb_mlp_hidden_fwd(_ptr(emb), _ptr(w1s), emb.shape[0], emb.shape[1], w1s.shape[1], _ptr(h),
                                               _ptr(h_lo), _stream()), "nqb_mlp_hidden_fwd")


def mlp_hidden_variant(variant: int = 0) -> int:
    """Kernel generation of ``mlp_hidden_fwd/bwd``: 2 = batched kernels, 1 = round-1 kernels; 0 only queries.
    Returns the previous value (``nqb_mlp_hidden_set_variant``; A/B timing and the variant parity test)."""
    return int(_capi.lib().nqb_mlp_hidden_set_variant(int(variant)))


def mlp_hidden_bwd(emb: torch.Tensor, w1s: torch.Tensor, gh: torch.Tensor, gemb: torch.Tensor) -> None:
    """``gemb = (gh * silu'(emb @ w1s)) @ w1s^T``."""
    _require_cuda(emb, w1s, gh, gemb)
    _capi.check(_capi.lib().nqb_mlp_hidden_bwd(_ptr(emb), _ptr(w1s), _ptr(gh), emb.shape[0], emb.shape[1], w1s.shape[1],
                                               _ptr(gemb), _stream()), "nqb_mlp_hidden_bwd")


# ---------------------------------------------------------------------------------------
# Gate nonlinearity -- nqb_gate_fwd / nqb_gate_bwd
# ---------------------------------------------------------------------------------------
class GateTables:
    """Column tables of the fused Gate kernels for ``irreps_in = scalars + gates + gated`` in ``layout``
    (e3nn ``nn.Gate``, nequip/nn/convnetlayer.py:104-112).  ``p`` of a scalar/gate irrep picks the activation:
    even -> c_silu * silu, odd -> c_tanh * tanh."""

    def __init__(self, irreps_scalars, irreps_gates, irreps_gated, layout: str, device):
        from .irreps import Irreps

        sc, ga, gd = Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated)
        if sum(m for m, _ in ga) != sum(m for m, _ in gd):
            raise ValueError("Gate: one gate per gated multiplicity")
        ns, ng = sc.dim, ga.dim
        self.d_in = ns + ng + gd.dim
        self.d_out = ns + gd.dim
        src, gate, kind = [0] * self.d_out, [-1] * self.d_out, [0] * self.d_out
        tab = [[0] * 6 for _ in range(self.d_in)]
        off = 0
        for mul, ir in sc:  # scalars: out[j] = act(x[j])
            k = 0 if ir.p == 1 else 1
            for u in range(mul):
                src[off + u], kind[off + u] = off + u, k
                tab[off + u] = [0, off + u, 0, 0, 0, k]
            off += mul
        gate_kind = []
        for mul, ir in ga:
            gate_kind += [0 if ir.p == 1 else 1] * mul
        g0, in_off, out_off = 0, ns + ng, ns
        for mul, ir in gd:  # gated chunk: out = x * act(gate of its multiplicity index u)
            d = ir.dim
            stride = mul if layout == "ir_mul" else 1  # distance between the 2l+1 components of one u
            for u in range(mul):
                gcol, k = ns + g0 + u, gate_kind[g0 + u]
                first = u if layout == "ir_mul" else u * d
                tab[gcol] = [2, out_off + first, in_off + first, stride, d, k]
                for c in range(d):
                    pos = first + c * stride
                    src[out_off + pos], gate[out_off + pos], kind[out_off + pos] = in_off + pos, gcol, k
                    tab[in_off + pos] = [1, out_off + pos, gcol, 0, 0, k]
            g0 += mul
            in_off += mul * d
            out_off += mul * d
        mk = lambda v: torch.tensor(v, dtype=torch.int32, device=device).contiguous()
        self.src, self.gate, self.kind = mk(src), mk(gate), mk(kind)
        self.tab = mk([x for row in tab for x in row])


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tabs: GateTables):
        x = x.contiguous()
        out = torch.empty((x.shape[0], tabs.d_out), dtype=x.dtype, device=x.device)
        _capi.check(_capi.lib().nqb_gate_fwd(_DT[x.dtype], _ptr(x), x.shape[0], tabs.d_in, tabs.d_out, _ptr(tabs.src),
                                             _ptr(tabs.gate), _ptr(tabs.kind), _ptr(out), _stream()), "nqb_gate_fwd")
        ctx.tabs = tabs
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        gout = gout.contiguous()
        gx = torch.empty_like(x)
        t = ctx.tabs
        _capi.check(_capi.lib().nqb_gate_bwd(_DT[x.dtype], _ptr(x), _ptr(gout), x.shape[0], t.d_in, t.d_out, _ptr(t.tab),
                                             _ptr(gx), _stream()), "nqb_gate_bwd")
        return gx, None


def gate(x: torch.Tensor, tabs: GateTables) -> torch.Tensor:
    """Fused Gate nonlinearity on ``x [N, scalars + gates + gated]`` (CUDA, float32/float64)."""
    _require_cuda(x)
    if x.dtype not in _DT or x.dim() != 2 or x.shape[1] != tabs.d_in:
        raise ValueError("gate: x must be [N, %d] float32/float64" % tabs.d_in)
    return _GateFn.apply(x, tabs)


# ---------------------------------------------------------------------------------------
# EXPERIMENTAL: transposed K <= 128 GEMM with TMEM-resident weights -- nqb_gemm_t_* (opt-in tests only)
# ---------------------------------------------------------------------------------------
class GemmT:
    """``C[M, N] = A[M, K] @ B[K, N]`` for K <= 128 with ``B`` prepared once."""

    def __init__(self, B: torch.Tensor, device, scale: float = 1.0):
        L = _capi.lib()
        self.K, self.N = int(B.shape[0]), int(B.shape[1])
        if self.K > 128 or self.K % 4:
            raise ValueError("GemmT: K must be a multiple of 4 and <= 128")
        Bc = B.detach().to(device=device, dtype=torch.float32).contiguous()
        self.prepared = torch.empty(int(L.nqb_gemm_t_prepared_floats(self.K, self.N)), dtype=torch.float32, device=device)
        _capi.check(L.nqb_gemm_t_prepare(_ptr(Bc), Bc.shape[1], self.K, self.N, 0, float(scale), _ptr(self.prepared),
                                         _stream()), "nqb_gemm_t_prepare")

    def run(self, a: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
        _require_cuda(a, c)
        if a.dtype != torch.float32 or c.dtype != torch.float32 or a.stride(1) != 1 or c.stride(1) != 1:
            raise TypeError("GemmT.run: float32 row-major only")
        M = a.shape[0]
        _capi.check(_capi.lib().nqb_gemm_t_run(_ptr(self.prepared), self.K, self.N, _ptr(a), a.stride(0), _ptr(c),
                                               c.stride(0), M, _stream()), "nqb_gemm_t_run")
        return c


# ---------------------------------------------------------------------------------------
# neighbour list on the device -- nqb_nl_bin / nqb_nl_count / nqb_nl_fill   
# ---------------------------------------------------------------------------------------
def neighbor_list(pos: torch.Tensor, cell=None, pbc=True, r_max: float = 5.0, transpose_perm: bool = False):
    """Full neighbour list within ``r_max`` built on the GPU (cell list), in the layout the convolution wants.

    ``pos`` [N,3] float64 CUDA; ``cell`` [3,3] (rows = lattice vectors; host or device) or None; ``pbc`` bool or 3
    bools.  Returns a dict with ``edge_index`` [2,E] int64 (row 0 = centre / scatter destination, row 1 =
    neighbour), ``edge_cell_shift`` [E,3] float64 (edge vector = pos[j] - pos[i] + shift @ cell), ``row_ptr``
    [N+1] int64 (destination CSR: edges are sorted by (centre, neighbour)) and, on request,
    ``edge_transpose_perm`` [E] (argsort by (neighbour, centre), nequip/data/transforms/neighborlist.py:150-155).
    Same contract as the reference's host backends (nequip/data/_nl.py:60-152): both directions, no self edge in
    the home image.  One host synchronisation (the edge count)."""
    import numpy as np

    _require_cuda(pos)
    L = _capi.lib()
    pos = pos.detach().double().contiguous()
    N = pos.shape[0]
    dev = pos.device
    if isinstance(pbc, bool):
        pbc = (pbc,) * 3
    pbc = [bool(b) for b in (pbc.tolist() if torch.is_tensor(pbc) else pbc)]
    if cell is None:
        if any(pbc):
            raise ValueError("Periodic boundary conditions requested but no cell was provided.")
        cell_np = np.eye(3)
    else:
        cell_np = (cell.detach().cpu().double().reshape(3, 3).numpy() if torch.is_tensor(cell) else np.asarray(cell, dtype=np.float64).reshape(3, 3)).copy()
    inv_np = np.linalg.inv(cell_np)
    # distance between opposite faces along each lattice direction = 1 / |column d of the inverse|
    perp = 1.0 / np.linalg.norm(inv_np, axis=0)
    lo = np.zeros(3)
    width = np.ones(3)
    if not all(pbc) and N > 0:
        frac = pos @ torch.as_tensor(inv_np, device=dev)
        fmin, fmax = frac.min(0).values.cpu().numpy(), frac.max(0).values.cpu().numpy()
        for d in range(3):
            if not pbc[d]:
                lo[d], width[d] = fmin[d], max(fmax[d] - fmin[d], 1e-9) * (1 + 1e-9)
    nb, sr = [1, 1, 1], [1, 1, 1]
    cap = max(1, int(round((4 * max(N, 1)) ** (1.0 / 3.0))))
    for d in range(3):
        extent = perp[d] * (1.0 if pbc[d] else width[d])
        nb[d] = int(min(cap, max(1, np.floor(extent / r_max))))
        sr[d] = int(np.ceil(r_max / (extent / nb[d]) - 1e-12)) if pbc[d] else 1
        sr[d] = max(sr[d], 1)
    I3 = C.c_int * 3
    D9, D3 = C.c_double * 9, C.c_double * 3
    cell_c, inv_c = D9(*cell_np.reshape(-1)), D9(*inv_np.reshape(-1))
    pbc_c, nb_c, sr_c = I3(*[int(b) for b in pbc]), I3(*nb), I3(*sr)
    lo_c, wd_c = D3(*lo), D3(*width)
    st = _stream()
    wpos = torch.empty((N, 3), dtype=torch.float64, device=dev)
    base = torch.empty((N, 3), dtype=torch.int32, device=dev)
    cidx = torch.empty((N, 3), dtype=torch.int32, device=dev)
    binid = torch.empty((N,), dtype=torch.int64, device=dev)
    _capi.check(L.nqb_nl_bin(_ptr(pos), N, cell_c, inv_c, pbc_c, nb_c, sr_c, lo_c, wd_c, float(r_max), _ptr(wpos), _ptr(base),
                             _ptr(binid), _ptr(cidx), st), "nqb_nl_bin")
    nbins = nb[0] * nb[1] * nb[2]
    sorted_bin, order = torch.sort(binid, stable=True)
    bin_start = torch.searchsorted(sorted_bin, torch.arange(nbins + 1, device=dev, dtype=torch.int64)).contiguous()
    counts = torch.zeros((N,), dtype=torch.int64, device=dev)
    _capi.check(L.nqb_nl_count(N, cell_c, inv_c, pbc_c, nb_c, sr_c, float(r_max), _ptr(wpos), _ptr(cidx), _ptr(order),
                               _ptr(bin_start), _ptr(counts), st), "nqb_nl_count")
    row_ptr = torch.zeros((N + 1,), dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    E = int(row_ptr[-1].item())
    edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
    shifts = torch.empty((E, 3), dtype=torch.float64, device=dev)
    _capi.check(L.nqb_nl_fill(N, E, cell_c, inv_c, pbc_c, nb_c, sr_c, float(r_max), _ptr(wpos), _ptr(cidx), _ptr(base),
                              _ptr(order), _ptr(bin_start), _ptr(row_ptr), _ptr(edge_index), _ptr(shifts), st), "nqb_nl_fill")
    out = {"edge_index": edge_index, "edge_cell_shift": shifts, "row_ptr": row_ptr}
    if transpose_perm:
        out["edge_transpose_perm"] = torch.argsort(edge_index[1] * N + edge_index[0], stable=True)
    return out
