"""Minimal irreps bookkeeping for the oracle (TEST INFRASTRUCTURE ONLY).

Restates the subset of ``e3nn.o3.Irreps`` semantics that the reference's hot path
relies on (call sites: ``nequip/nn/interaction_block.py:89-109``,
``nequip/nn/convnetlayer.py:74-114``, ``nequip/model/nequip_models.py:160-200``):

* an irrep is ``(l, p)`` with ``p in {+1 (e), -1 (o)}``; dim ``2l+1``
* irreps are ordered lists of ``(mul, (l, p))``; data layout is **mul_ir**
  (each chunk is ``[mul, 2l+1]`` row-major)
* ``ir1 * ir2`` enumerates ``l = |l1-l2| .. l1+l2`` with parity ``p1*p2``
* ``sort()`` is a stable sort by ``(l, p)`` -- e3nn's ``Irrep`` is a
  ``(l, p)`` tuple so ``1o=(1,-1)`` sorts before ``1e=(1,+1)``
* ``simplify()`` merges adjacent equal irreps
"""
import re
from typing import List, Tuple

Ir = Tuple[int, int]  # (l, p)
IrList = List[Tuple[int, Ir]]  # [(mul, (l, p)), ...]

_TERM = re.compile(r"^\s*(?:(\d+)\s*x\s*)?(\d+)\s*([eo])\s*$")


def parse(s) -> IrList:
    """``"32x0e + 32x1o"`` -> ``[(32,(0,1)), (32,(1,-1))]``; lists pass through."""
    if not isinstance(s, str):
        return [(int(m), (int(ir[0]), int(ir[1]))) for m, ir in s]
    out = []
    s = s.strip()
    if not s:
        return out
    for term in s.split("+"):
        m = _TERM.match(term)
        if m is None:
            raise ValueError(f"cannot parse irreps term {term!r}")
        mul = int(m.group(1)) if m.group(1) else 1
        out.append((mul, (int(m.group(2)), 1 if m.group(3) == "e" else -1)))
    return out


def fmt(irreps: IrList) -> str:
    return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, (l, p) in irreps)


def ir_dim(ir: Ir) -> int:
    return 2 * ir[0] + 1


def dim(irreps: IrList) -> int:
    return sum(m * ir_dim(ir) for m, ir in irreps)


def num_irreps(irreps: IrList) -> int:
    return sum(m for m, _ in irreps)


def slices(irreps: IrList) -> List[slice]:
    out, off = [], 0
    for m, ir in irreps:
        n = m * ir_dim(ir)
        out.append(slice(off, off + n))
        off += n
    return out


def ir_mul(ir1: Ir, ir2: Ir) -> List[Ir]:
    (l1, p1), (l2, p2) = ir1, ir2
    return [(l, p1 * p2) for l in range(abs(l1 - l2), l1 + l2 + 1)]


def contains(irreps: IrList, ir: Ir) -> bool:
    return any(ir == ir_ for _, ir_ in irreps)


def sort(irreps: IrList):
    """Stable sort by (l, p).  Returns ``(sorted, p, inv)`` with e3nn's meaning:
    ``p[i_old] = i_new`` and ``inv[i_new] = i_old``."""
    inv = sorted(range(len(irreps)), key=lambda i: (irreps[i][1], i))
    p = [0] * len(irreps)
    for new, old in enumerate(inv):
        p[old] = new
    return [irreps[i] for i in inv], p, inv


def simplify(irreps: IrList) -> IrList:
    out: IrList = []
    for m, ir in irreps:
        if m == 0:
            continue
        if out and out[-1][1] == ir:
            out[-1] = (out[-1][0] + m, ir)
        else:
            out.append((m, ir))
    return out


def spherical_harmonics(lmax: int, p: int = -1) -> IrList:
    """``Irreps.spherical_harmonics(lmax)``: ``1x l`` with parity ``p**l``."""
    return [(1, (l, p**l)) for l in range(lmax + 1)]


def build_tp_instructions(feature_irreps_in, irreps_edge_attr, feature_irreps_out):
    """Path enumeration of ``InteractionBlock.__init__``
    (``nequip/nn/interaction_block.py:89-109``): returns (irreps_mid_sorted,
    instructions) where instructions are ``(i_in1, i_in2, i_out, "uvu", True)``
    in *instruction-list order* (weights are flattened in this order) with
    ``i_out`` already permuted to the sorted ``irreps_mid``."""
    fin, fe, fout = parse(feature_irreps_in), parse(irreps_edge_attr), parse(feature_irreps_out)
    mid, ins = [], []
    for i, (mul, ir_in) in enumerate(fin):
        for j, (_, ir_e) in enumerate(fe):
            for ir_out in ir_mul(ir_in, ir_e):
                if contains(fout, ir_out):
                    k = len(mid)
                    mid.append((mul, ir_out))
                    ins.append((i, j, k, "uvu", True))
    mid_sorted, p, _ = sort(mid)
    ins = [(a, b, p[c], mode, tr) for a, b, c, mode, tr in ins]
    return mid_sorted, ins
