"""OPT-IN (``NQB_EXPERIMENTAL=1``): the round-2 candidate GEMM with TMEM-resident weights
(nequip_b200/csrc/nqb_gemm_t.cu).  Not part of the default GPU suite: it has never run on hardware yet."""
import os

import pytest
import torch

from nequip_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("NQB_EXPERIMENTAL") != "1",
                                                  reason="experimental kernel: set NQB_EXPERIMENTAL=1")]


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,K,N", [(64, 8, 128), (100, 128, 128), (1000, 128, 864), (4099, 64, 200), (70000, 128, 1728)])
def test_gemm_t_matches_float64(M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(K, N, generator=g)
    gt = ops.GemmT(B, "cuda", scale=0.5)
    C = torch.full((M, N), float("nan"), device="cuda")
    gt.run(A, C)
    torch.cuda.synchronize()
    ref = A.double().cpu() @ B.double() * 0.5
    err = (C.cpu().double() - ref).abs().max().item()
    assert err <= 1.5e-6 * ref.abs().max().item() + 1e-7, err
