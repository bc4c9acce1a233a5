#!/usr/bin/env python
"""Generate the committed fixtures tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).

The reference cannot be imported here (e3nn is absent) and its tests hold no golden
vectors for this path, so these files do NOT pin the oracle to the reference; they freeze the numbers
the oracle produced when its mathematical identities were verified (tests/test_oracle_math.py), so
that later edits to the oracle or to the kernels cannot drift unnoticed.  Everything is float64 and
seeded; the files are a few tens of KB."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import irreps as I  # noqa: E402
from oracle import sh as osh  # noqa: E402
from oracle import tp as otp  # noqa: E402
from oracle import wigner as ow  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# the irreps grid of the reference's kernel test (tests/unit/nn/test_tp_scatter_kernel.py:12-31)
TP_CASES = [
    ("4x0e+4x1o", "1x0e+1x1o", "4x0e+4x1o+4x1e"),
    ("8x0e+8x1o+8x2e", "1x0e+1x1o+1x2e", "8x0e+8x1o+8x2e"),
    ("4x0e+4x0o+4x1o+4x1e+4x2e+4x2o", "1x0e+1x1o+1x2e+1x3o", "4x0e+4x1o+4x2e+4x3o"),
]


def main():
    # 1. real Wigner-3j blocks up to l = 3
    w3j = {}
    for l1 in range(4):
        for l2 in range(4):
            for l3 in range(abs(l1 - l2), min(3, l1 + l2) + 1):
                w3j[f"w3j_{l1}{l2}{l3}"] = ow.wigner_3j(l1, l2, l3)
    np.savez_compressed(os.path.join(OUT, "w3j_lmax3.npz"), **w3j)

    # 2. component-normalised spherical harmonics of 16 fixed directions (non-unit vectors: normalize=True)
    g = torch.Generator().manual_seed(11)
    vec = torch.randn(16, 3, generator=g, dtype=torch.float64) * 2.0
    np.savez_compressed(os.path.join(OUT, "sh_lmax3.npz"), vec=vec.numpy(),
                        y=osh.spherical_harmonics(3, vec, normalize=True).numpy())

    # 3. TensorProductScatter forward + the three gradients on the reference's test grid (N=8, E=15)
    out = {}
    for ci, (fin, fe, fout) in enumerate(TP_CASES):
        mid, ins = I.build_tp_instructions(I.parse(fin), I.parse(fe), I.parse(fout))
        g = torch.Generator().manual_seed(100 + ci)
        N, E = 8, 15
        x = torch.randn(N, I.dim(I.parse(fin)), generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.randn(E, I.dim(I.parse(fe)), generator=g, dtype=torch.float64, requires_grad=True)
        w = torch.randn(E, otp.weight_numel(I.parse(fin), I.parse(fe), ins), generator=g, dtype=torch.float64,
                        requires_grad=True)
        dst = torch.randint(0, N, (E,), generator=g)
        src = torch.randint(0, N, (E,), generator=g)
        o = otp.tp_scatter(x, y, w, dst, src, I.parse(fin), I.parse(fe), mid, ins)
        go = torch.randn(o.shape, generator=g, dtype=torch.float64)
        gx, gy, gw = torch.autograd.grad([o], [x, y, w], [go])
        for k, v in dict(x=x, y=y, w=w, dst=dst, src=src, out=o, go=go, gx=gx, gy=gy, gw=gw).items():
            out[f"c{ci}_{k}"] = v.detach().numpy()
        out[f"c{ci}_irreps"] = np.array([fin, fe, fout, I.fmt(mid)])
        out[f"c{ci}_instructions"] = np.array([[a, b, c] for (a, b, c, _m, _t) in ins], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "tp_scatter_grid.npz"), **out)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
