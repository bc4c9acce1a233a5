#!/usr/bin/env python
"""Time the experimental TMEM-resident-weights GEMM (nqb_gemm_t.cu) on the radial-MLP forward shapes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nequip_b200 import ops  # noqa: E402

E = 588616
for (name, K, N) in [("mlp_fwd_L0", 128, 192), ("mlp_fwd_L1", 128, 960), ("mlp_fwd_L2", 128, 1728)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(E, K, device="cuda", generator=g)
    B = torch.randn(K, N, device="cuda", generator=g)
    C = torch.empty(E, N, device="cuda")
    gt = ops.GemmT(B, "cuda")
    for _ in range(2):
        gt.run(A, C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        gt.run(A, C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    ref = A[:4096].double() @ B.double()
    err = float((C[:4096].double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({"case": name, "ms": round(ms, 4), "TFLOPs_fp32_equiv": round(2.0 * E * K * N / ms / 1e9, 1), "rel_err": err}), flush=True)
    del A, B, C
