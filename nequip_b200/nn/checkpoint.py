"""State-dict key mapping between ``NequIPEnergyModel`` and the reference's ``NequIPGNNEnergyModel``.

The reference assembles a ``SequentialGraphNetwork`` with the module names of
nequip/model/nequip_models.py:288-399 -- ``type_embed`` (``NodeTypeEmbed.embed_module``, nn/embedding/node.py:75),
``layer{i}_convnet.conv.{linear_1, linear_2, sc, edge_mlp.mlp.{2k}}`` (nn/convnetlayer.py:142,
nn/interaction_block.py:82-146, nn/mlp.py:134-192), ``per_atom_energy_readout.mlp_module`` (nn/mlp.py:62),
``per_type_energy_scale_shift.{scales, shifts}`` (nn/atomwise.py:206-233) -- wrapped in ``ForceStressOutput.func``
and possibly ``GraphModel.model``.  Parameters are matched by SUFFIX, so any wrapper prefix is accepted; e3nn's
persistent buffers (``tp_scatter.tp.*``, ``*.output_mask``, ...) carry no learnable state and are ignored.

Flattening conventions assumed for the e3nn weights (e3nn itself is not installed here, so
this is stated, not verified): ``o3.Linear.weight`` = concatenation over instructions (i_in-major over equal irreps)
of row-major ``[mul_in, mul_out]`` blocks; ``FullyConnectedTensorProduct.weight`` = concatenation of ``[mul_in,
num_attr, mul_out]`` blocks; ``ScalarLinearLayer.weight`` = ``[in, out]``.
"""
from __future__ import annotations

import re
from typing import Dict, Tuple

import torch

_IGNORED = re.compile(r"(tp_scatter\.tp\.|\.output_mask$|_w3j|\.tp\._|norm_const$|\.alpha$|bessel_weights$|_empty$)")


def reference_key_map(num_layers: int, radial_mlp_depth: int = 1) -> Dict[str, str]:
    """reference key suffix -> ``NequIPEnergyModel.state_dict()`` key."""
    m = {
        "type_embed.embed_module.weight": "type_embed.weight",
        "per_atom_energy_readout.mlp_module.mlp.0.weight": "readout.mlp.0.weight",
        "per_type_energy_scale_shift.scales": "scales",
        "per_type_energy_scale_shift.shifts": "shifts",
    }
    for i in range(num_layers):
        ref, ours = f"layer{i}_convnet.conv.", f"layers.{i}.conv."
        m[ref + "linear_1.weight"] = ours + "linear_1.weight"
        m[ref + "linear_2.weight"] = ours + "linear_2.weight"
        if i != 0:
            m[ref + "sc.weight"] = ours + "sc.weight"
        for q in range(radial_mlp_depth + 1):
            m[ref + f"edge_mlp.mlp.{2 * q}.weight"] = ours + f"edge_mlp.mlp.{2 * q}.weight"
    return m


def to_reference_state_dict(model, prefix: str = "model.func.") -> Dict[str, torch.Tensor]:
    """This model's parameters under the reference's names (e3nn buffers are not produced)."""
    cfg = model.config
    inv = {v: k for k, v in reference_key_map(cfg["num_layers"], cfg["radial_mlp_depth"]).items()}
    out = {}
    for k, v in model.state_dict().items():
        if k in inv:
            if k in ("scales", "shifts") and v.numel() == 0:
                continue
            out[prefix + inv[k]] = v.detach().clone()
    return out


def load_reference_state_dict(model, ref_sd: Dict[str, torch.Tensor], strict: bool = True) -> Tuple[list, list]:
    """Load a reference (nequip) state dict into ``model``.  Returns (missing, unexpected) like
    ``torch.nn.Module.load_state_dict``; with ``strict`` both must be empty (ignored e3nn buffers aside)."""
    cfg = model.config
    kmap = reference_key_map(cfg["num_layers"], cfg["radial_mlp_depth"])
    own = model.state_dict()
    new, unexpected, used = {}, [], set()
    for rk, v in ref_sd.items():
        hit = [s for s in kmap if rk == s or rk.endswith("." + s)]
        if not hit:
            if not _IGNORED.search(rk):
                unexpected.append(rk)
            continue
        ok = kmap[max(hit, key=len)]
        t = own[ok]
        if ok in ("scales", "shifts"):
            v = v.reshape(-1, 1).to(torch.float64)
            if v.numel() == 1 and t.numel() > 1:
                v = v.expand_as(t).clone()
            if t.numel() == 0:  # the model was built without scale / shift: adopt the checkpoint's table
                getattr(model, ok).resize_(v.shape)
                t = getattr(model, ok)
        if tuple(v.shape) != tuple(t.shape):
            raise ValueError(f"load_reference_state_dict: {rk} has shape {tuple(v.shape)}, expected {tuple(t.shape)} ({ok})")
        new[ok] = v.to(t.dtype)
        used.add(ok)
    missing = [k for k in own if k not in used and not (k in ("scales", "shifts") and own[k].numel() == 0)]
    if strict and (missing or unexpected):
        raise KeyError(f"load_reference_state_dict: missing {missing}, unexpected {unexpected}")
    merged = dict(model.state_dict())
    merged.update(new)
    model.load_state_dict(merged)
    return missing, unexpected
