"""Signatures whose kernel libraries are prebuilt by ``__graft_entry__.build()`` so that
they travel to the GPU box with the snapshot (anything else is compiled on first use).
"""
from __future__ import annotations

from typing import List, Tuple

from .codegen import TPSignature
from .irreps import Irreps, build_tp_instructions

# the reference's kernel test grid (tests/unit/nn/test_tp_scatter_kernel.py:38-55)
TEST_FEATURE_IRREPS_IN = ["4x0e + 3x1o + 2x2e", "2x0e + 2x1o + 2x2e", "8x0e + 8x2e + 8x1o"]
TEST_IRREPS_EDGE_ATTR = ["0e + 1o", "0e + 1o + 2e"]
TEST_IRREPS_MID = ["0e + 1o + 2e", "2x0e + 2x1o + 2x2e", "24x0e + 32x1o + 16x1e + 16x2o + 32x2e"]


def make_signature(feature_irreps_in, irreps_edge_attr, feature_irreps_out) -> TPSignature:
    """Signature exactly as ``InteractionBlock`` would build it (interaction_block.py:89-116)."""
    mid, ins = build_tp_instructions(feature_irreps_in, irreps_edge_attr, feature_irreps_out)
    return TPSignature(Irreps(feature_irreps_in), Irreps(irreps_edge_attr), mid, ins)


def reference_test_grid() -> List[TPSignature]:
    out = []
    for fin in TEST_FEATURE_IRREPS_IN:
        for fe in TEST_IRREPS_EDGE_ATTR:
            for fm in TEST_IRREPS_MID:
                try:
                    out.append(make_signature(fin, fe, fm))
                except ValueError:
                    pass  # no valid instruction (the reference test skips these)
    return out


def nequip_layer_signatures(l_max: int, num_features: int, num_layers: int, parity: bool = True) -> List[TPSignature]:
    """Per-layer signatures of ``NequIPGNNModel`` (nequip/model/nequip_models.py:116-210 +
    nequip/nn/convnetlayer.py:74-114): returns one TPSignature per interaction layer."""
    from .nn.model import layer_irreps  # local import: nn.model imports this package

    return [make_signature(fin, fe, fout) for (fin, fe, fout, _gate) in layer_irreps(l_max, num_features, num_layers, parity)]


def all_known() -> List[TPSignature]:
    sigs = reference_test_grid()
    for (lm, nf, nl) in [(1, 32, 4), (2, 32, 4), (2, 64, 4), (3, 32, 5), (2, 8, 3), (1, 8, 2)]:
        sigs += nequip_layer_signatures(lm, nf, nl)
    uniq = {}
    for s in sigs:
        uniq[s.canonical()] = s
    return list(uniq.values())
