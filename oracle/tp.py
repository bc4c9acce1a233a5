"""e3nn-formulation tensor product + scatter on CPU (TEST INFRASTRUCTURE ONLY).

Restates, in the reference's own formulation (gather ``x[edge_src]``, one
einsum with the real w3j per path, per-edge weights, concatenate ``[E, D_mid]``,
then ``zeros.scatter_add_``):

* ``TensorProductScatter.forward``   nequip/nn/_tp_scatter_base.py:35-38
* ``o3.TensorProduct(..., shared_weights=False, internal_weights=False)``
  constructed at nequip/nn/_tp_scatter_base.py:24-31 (math in e3nn 0.6.x,
  ``irrep_normalization="component"``, ``path_normalization="element"``)
* ``scatter``                        nequip/nn/utils.py:24-53

Differentiable through torch autograd (used as the gradient oracle too).
"""
import math
from typing import List, Sequence, Tuple

import torch

from . import irreps as I
from . import wigner


def path_coefficients(irreps_in1, irreps_in2, irreps_out, instructions) -> List[float]:
    """sqrt(alpha) per instruction; e3nn ``TensorProduct.__init__`` with
    component irrep normalisation and element path normalisation, all
    variances 1, path_weight 1.  For ``uvu``: num_elements = mul_in2."""
    in1, in2, out = I.parse(irreps_in1), I.parse(irreps_in2), I.parse(irreps_out)

    def num_elements(ins):
        mode = ins[3]
        m1, m2 = in1[ins[0]][0], in2[ins[1]][0]
        return {"uvw": m1 * m2, "uvu": m2, "uvv": m1, "uuw": m1, "uuu": 1, "uvuv": 1}[mode]

    coeffs = []
    for ins in instructions:
        x = sum(num_elements(j) for j in instructions if j[2] == ins[2])
        alpha = I.ir_dim(out[ins[2]][1]) / x
        coeffs.append(math.sqrt(alpha))
    return coeffs


def weight_numel(irreps_in1, irreps_in2, instructions) -> int:
    in1, in2 = I.parse(irreps_in1), I.parse(irreps_in2)
    return sum(in1[i][0] * in2[j][0] for i, j, _, mode, hw in instructions if hw)


def tensor_product_uvu(
    x1: torch.Tensor,
    x2: torch.Tensor,
    weight: torch.Tensor,
    irreps_in1,
    irreps_in2,
    irreps_out,
    instructions: Sequence[Tuple],
) -> torch.Tensor:
    """``[E, D_in1] x [E, D_in2] x [E, W] -> [E, D_out]`` (all mul_ir layout).

    out[z, u, k] = sqrt(alpha) * sum_v w[z,u,v] sum_ij C[i,j,k] x1[z,u,i] x2[z,v,j]
    """
    in1, in2, out = I.parse(irreps_in1), I.parse(irreps_in2), I.parse(irreps_out)
    s1, s2, so = I.slices(in1), I.slices(in2), I.slices(out)
    coeffs = path_coefficients(in1, in2, out, instructions)
    E = x1.shape[0]
    out_chunks = [None] * len(out)
    woff = 0
    for ins, c in zip(instructions, coeffs):
        i1, i2, io, mode, has_w = ins
        assert mode == "uvu" and has_w
        m1, (l1, p1) = in1[i1]
        m2, (l2, p2) = in2[i2]
        mo, (l3, p3) = out[io]
        assert mo == m1 and p3 == p1 * p2 and abs(l1 - l2) <= l3 <= l1 + l2
        C = torch.from_numpy(wigner.wigner_3j(l1, l2, l3).copy()).to(x1.dtype)
        a = x1[:, s1[i1]].reshape(E, m1, 2 * l1 + 1)
        b = x2[:, s2[i2]].reshape(E, m2, 2 * l2 + 1)
        w = weight[:, woff : woff + m1 * m2].reshape(E, m1, m2)
        woff += m1 * m2
        if m2 == 1:
            # e3nn's generated code for 'uvu': outer product xx = x1 (x) x2, contract with the
            # w3j as one dense matmul, then scale by the per-edge weight
            xx = (a.unsqueeze(-1) * b.reshape(E, 1, 1, 2 * l2 + 1)).reshape(E * m1, (2 * l1 + 1) * (2 * l2 + 1))
            r = torch.mm(xx, C.reshape((2 * l1 + 1) * (2 * l2 + 1), 2 * l3 + 1)).reshape(E, m1, 2 * l3 + 1)
            r = c * r * w.reshape(E, m1, 1)
        else:
            r = c * torch.einsum("ijk,zuv,zui,zvj->zuk", C, w, a, b)
        r = r.reshape(E, m1 * (2 * l3 + 1))
        out_chunks[io] = r if out_chunks[io] is None else out_chunks[io] + r
    assert woff == weight.shape[1]
    for io, ch in enumerate(out_chunks):
        if ch is None:
            out_chunks[io] = x1.new_zeros(E, out[io][0] * I.ir_dim(out[io][1]))
    return torch.cat(out_chunks, dim=1)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``nequip.nn.utils.scatter`` (dim=0, reduce="sum")."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def tp_scatter(
    x, edge_attr, edge_weight, edge_dst, edge_src,
    feature_irreps_in, irreps_edge_attr, irreps_mid, instructions,
    chunk: int = 0,
):
    """``TensorProductScatter.forward``.  ``chunk>0`` evaluates edges in chunks
    (identical math, bounded memory) for the large CPU-baseline cases."""
    N, E = x.shape[0], edge_dst.shape[0]
    if chunk <= 0 or E <= chunk:
        ef = tensor_product_uvu(x[edge_src], edge_attr, edge_weight,
                                feature_irreps_in, irreps_edge_attr, irreps_mid, instructions)
        return scatter_sum(ef, edge_dst, N)
    out = torch.zeros(N, I.dim(I.parse(irreps_mid)), dtype=x.dtype)
    for s in range(0, E, chunk):
        sl = slice(s, min(E, s + chunk))
        ef = tensor_product_uvu(x[edge_src[sl]], edge_attr[sl], edge_weight[sl],
                                feature_irreps_in, irreps_edge_attr, irreps_mid, instructions)
        out = out.index_add(0, edge_dst[sl], ef)
    return out
