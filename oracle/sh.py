"""Real spherical harmonics of edge vectors (TEST INFRASTRUCTURE ONLY).

Restates ``e3nn.o3.SphericalHarmonics(irreps, normalize=True,
normalization="component")`` as constructed by the reference at
``nequip/nn/embedding/_edge.py:187-189`` and evaluated at ``:193-198``
(fp64 in, cast to model dtype by the caller).  e3nn 0.6.x is not vendored; the
closed forms below are e3nn's generated polynomials
(``e3nn/o3/_spherical_harmonics.py``: y is the polar axis, m ordered -l..l),
cross-checked in ``tests/test_oracle_math.py`` against the recurrence
``Y_{l+1} = normalise(C^{l+1,1,l} : Y_1 Y_l)`` built from ``oracle.wigner``.
"""
import math

import torch

from . import wigner


def sh_closed_form(lmax: int, vec: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """``[..., 3] -> [..., (lmax+1)^2]``, ``component`` normalisation
    (||Y_l||^2 = 2l+1 on the unit sphere).  lmax <= 3."""
    assert lmax <= 3
    if normalize:
        # torch.nn.functional.normalize semantics: x / max(||x||, eps)
        vec = vec / vec.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s3 = math.sqrt(3.0)
        out += [s3 * x, s3 * y, s3 * z]
    if lmax >= 2:
        s15, s5 = math.sqrt(15.0), math.sqrt(5.0)
        x2, y2, z2 = x * x, y * y, z * z
        sh20 = s15 * x * z
        sh21 = s15 * x * y
        sh22 = s5 * (y2 - 0.5 * (x2 + z2))
        sh23 = s15 * y * z
        sh24 = 0.5 * s15 * (z2 - x2)
        out += [sh20, sh21, sh22, sh23, sh24]
    if lmax >= 3:
        x2z2 = x2 + z2
        out += [
            (1.0 / 6.0) * math.sqrt(42.0) * (sh20 * z + sh24 * x),
            math.sqrt(7.0) * sh20 * y,
            (1.0 / 8.0) * math.sqrt(168.0) * (4.0 * y2 - x2z2) * x,
            0.5 * math.sqrt(7.0) * y * (2.0 * y2 - 3.0 * x2z2),
            (1.0 / 8.0) * math.sqrt(168.0) * z * (4.0 * y2 - x2z2),
            math.sqrt(7.0) * sh24 * y,
            (1.0 / 6.0) * math.sqrt(42.0) * (sh24 * z - sh20 * x),
        ]
    return torch.stack(out, dim=-1)


def sh_recurrence(lmax: int, vec: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """Any lmax, via Y_{l+1}[k] = c_l * sum_ab C^{l+1,1,l}[k,a,b] Y_1[a] Y_l[b],
    c_l > 0 fixed by ||Y_{l+1}||^2 = 2l+3 on the unit sphere."""
    if normalize:
        vec = vec / vec.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    y1 = math.sqrt(3.0) * vec
    blocks = [torch.ones_like(vec[..., :1])]
    if lmax >= 1:
        blocks.append(y1)
    # normalisation constants from a fixed probe direction
    probe = torch.tensor([0.3, -0.5, 0.8124038404635961], dtype=torch.float64)
    probe = probe / probe.norm()
    p_blocks = [torch.ones(1, dtype=torch.float64), math.sqrt(3.0) * probe]
    for l in range(1, lmax):
        C = torch.from_numpy(wigner.wigner_3j(l + 1, 1, l).copy())
        pn = torch.einsum("kab,a,b->k", C, p_blocks[1], p_blocks[l])
        c = math.sqrt(2 * l + 3) / float(pn.norm())
        p_blocks.append(c * pn)
        Cv = C.to(vec.dtype)
        nxt = c * torch.einsum("kab,...a,...b->...k", Cv, y1, blocks[l])
        blocks.append(nxt)
    return torch.cat(blocks[: lmax + 1], dim=-1)


def spherical_harmonics(lmax: int, vec: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    return sh_closed_form(lmax, vec, normalize) if lmax <= 3 else sh_recurrence(lmax, vec, normalize)
