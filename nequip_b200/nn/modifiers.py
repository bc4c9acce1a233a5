"""Model modifiers for the reference's seams next to ``TensorProductScatter``.

* ``B200SphericalHarmonicEdgeAttrs`` / ``enable_B200EdgeEmbed`` -- ``SphericalHarmonicEdgeAttrs.forward``
  (nequip/nn/embedding/_edge.py:153-198) on ``nqb_sh_fwd/bwd`` (fp64 evaluation, cast to the model dtype, exactly as
  :196-197 does);
* ``B200GhostExchangeModule`` / ``enable_B200GhostExchange`` -- a ``GhostExchangeModule``
  (nequip/nn/_ghost_exchange_base.py:8-57) whose exchange is ``nequip_b200.parallel.HaloExchange`` (NCCL
  all-to-all with the transposed exchange as backward), the counterpart of ``LAMMPSMLIAPGhostExchangeModule``
  (nequip/nn/_ghost_exchange_lmp_mliap.py:38-64) for ``torch.distributed`` hosts.  The exchange object travels in
  the data dict under ``NQB_HALO_KEY`` the way ``LMP_MLIAP_DATA_KEY`` does.

With nequip importable the classes subclass the reference modules and the modifiers are attached with the same
``@model_modifier`` mechanism as ``enable_OpenEquivariance`` (nequip/nn/_tp_scatter_base.py:40-77), so
``nequip.model.modify`` / ``nequip-compile --modifiers ...`` find them; without it (this container) they stand on
interface classes carrying the same attributes and are exercised by tests/test_modifiers_gpu.py.
"""
from __future__ import annotations

from typing import Dict

import torch

from .. import ops
from ..irreps import Irreps

NQB_HALO_KEY = "nqb_halo_exchange"
EDGE_VECTORS_KEY, EDGE_ATTRS_KEY, NODE_FEATURES_KEY = "edge_vectors", "edge_attrs", "node_features"

try:  # pragma: no cover - only where nequip + e3nn are installed
    from nequip.nn._ghost_exchange_base import GhostExchangeModule as _RefGhost, NoOpGhostExchangeModule as _RefNoOpGhost
    from nequip.nn.embedding._edge import SphericalHarmonicEdgeAttrs as _RefSH
    from nequip.nn.model_modifier_utils import model_modifier, replace_submodules
    from nequip.nn.utils import with_edge_vectors_ as _with_edge_vectors

    _HAVE_NEQUIP = True
except Exception:
    _RefGhost = _RefNoOpGhost = _RefSH = None
    _HAVE_NEQUIP = False


def _edge_vectors(data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``with_edge_vectors_`` (nequip/nn/utils.py:68-118): keep given vectors, else pos[j] - pos[i] + shift @ cell."""
    if EDGE_VECTORS_KEY in data:
        return data
    pos, ei = data["pos"], data["edge_index"]
    vec = torch.index_select(pos, 0, ei[1]) - torch.index_select(pos, 0, ei[0])
    if "cell" in data:
        vec = vec + torch.sum(data["edge_cell_shift"].to(pos.dtype).view(-1, 3, 1) * data["cell"].view(3, 3), 1)
    data[EDGE_VECTORS_KEY] = vec
    return data


class _SHInterface(torch.nn.Module):
    def __init__(self, irreps_edge_sh, edge_sh_normalization: str = "component", edge_sh_normalize: bool = True,
                 irreps_in=None, out_field: str = EDGE_ATTRS_KEY):
        super().__init__()
        self.out_field = out_field
        self.irreps_edge_sh = Irreps.spherical_harmonics(irreps_edge_sh) if isinstance(irreps_edge_sh, int) else Irreps(irreps_edge_sh)
        self.irreps_in = irreps_in
        self._output_dtype = torch.get_default_dtype()


class B200SphericalHarmonicEdgeAttrs(_RefSH if _HAVE_NEQUIP else _SHInterface):
    """``data[out_field] = Y_lm(edge_vectors)`` (component normalisation, unit vectors) on the CUDA kernel."""

    def __init__(self, irreps_edge_sh, edge_sh_normalization: str = "component", edge_sh_normalize: bool = True,
                 irreps_in=None, out_field: str = EDGE_ATTRS_KEY):
        if edge_sh_normalization != "component" or not edge_sh_normalize:
            raise NotImplementedError("B200SphericalHarmonicEdgeAttrs: only normalize=True, normalization='component' "
                                      "(what NequIPGNNModel builds, nequip/model/nequip_models.py:300-306)")
        super().__init__(irreps_edge_sh, edge_sh_normalization, edge_sh_normalize, irreps_in, out_field)
        ls = [ir.l for _, ir in Irreps(self.irreps_edge_sh)] if not _HAVE_NEQUIP else [ir.l for _, ir in self.irreps_edge_sh]
        if ls != list(range(len(ls))) or len(ls) > 4:
            raise NotImplementedError("B200SphericalHarmonicEdgeAttrs: irreps must be 0e + 1o + ... up to l = 3")
        self._lmax = len(ls) - 1

    def forward(self, data):
        data = _with_edge_vectors(data, with_lengths=False) if _HAVE_NEQUIP else _edge_vectors(data)
        data[self.out_field] = ops.spherical_harmonics(data[EDGE_VECTORS_KEY], self._lmax, self._output_dtype)
        return data


class _GhostInterface(torch.nn.Module):
    def __init__(self, field: str = NODE_FEATURES_KEY, irreps_in=None):
        super().__init__()
        self.field = field
        self.irreps_in = irreps_in or {}


class B200GhostExchangeModule(_RefGhost if _HAVE_NEQUIP else _GhostInterface):
    """Owned rows -> owned + ghost rows through ``data[NQB_HALO_KEY]`` (a ``nequip_b200.parallel.HaloExchange``)."""

    def forward(self, data, ghost_included: bool = False):
        if NQB_HALO_KEY not in data:
            raise RuntimeError("B200GhostExchangeModule needs data['%s'] (a nequip_b200.parallel.HaloExchange)" % NQB_HALO_KEY)
        halo = data[NQB_HALO_KEY]
        x = data[self.field]
        if ghost_included:
            x = torch.narrow(x, 0, 0, halo.plan.n_own)
        data[self.field] = halo(x)
        return data


def _swap(model: torch.nn.Module, target_cls, factory) -> torch.nn.Module:
    for name, child in list(model.named_children()):
        if isinstance(child, target_cls) and not isinstance(child, (B200SphericalHarmonicEdgeAttrs, B200GhostExchangeModule)):
            model._modules[name] = factory(child)
        else:
            _swap(child, target_cls, factory)
    return model


def _sh_factory(old):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(old._output_dtype)
    try:
        return B200SphericalHarmonicEdgeAttrs(old.irreps_edge_sh, irreps_in=getattr(old, "irreps_in", None), out_field=old.out_field)
    finally:
        torch.set_default_dtype(prev)


def enable_B200EdgeEmbed(model: torch.nn.Module) -> torch.nn.Module:
    """Swap every ``SphericalHarmonicEdgeAttrs`` for the CUDA kernel (stand-alone form of the modifier)."""
    return _swap(model, _RefSH if _HAVE_NEQUIP else _SHInterface, _sh_factory)


def enable_B200GhostExchange(model: torch.nn.Module) -> torch.nn.Module:
    """Swap every (no-op) ``GhostExchangeModule`` for the NCCL halo exchange."""
    return _swap(model, _RefNoOpGhost if _HAVE_NEQUIP else _GhostInterface,
                 lambda old: B200GhostExchangeModule(field=old.field, irreps_in=old.irreps_in))


if _HAVE_NEQUIP:  # pragma: no cover
    _RefSH.enable_B200EdgeEmbed = model_modifier(persistent=False, private=False, unsupported_devices=["cpu"],
                                                 supported_compile_modes=[])(
        classmethod(lambda cls, model: replace_submodules(model, cls, _sh_factory)))
    _RefNoOpGhost.enable_B200GhostExchange = model_modifier(persistent=True, private=True)(
        classmethod(lambda cls, model: replace_submodules(
            model, cls, lambda old: B200GhostExchangeModule(field=old.field, irreps_in=old.irreps_in))))
