"""fp32 accuracy of the tensor-core GEMM at the longest k the paper sweeps (d up to 262 144, NSDI'19 Fig. 5):
the fp32 fold of the TMEM accumulator every k_chunk keeps the error at the level of a sequential fp32 sum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dist", ["uniform01", "signed"])
def test_gemm_k_262144(ctx, dist):
    M, N, K = 512, 384, 262144
    gen = torch.Generator(device="cuda"); gen.manual_seed(11)
    A = torch.rand((M, K), device="cuda", generator=gen); B = torch.rand((K, N), device="cuda", generator=gen)
    if dist == "signed":
        A = A - 0.5; B = B - 0.5
    C = torch.full((M, N), float("nan"), device="cuda")
    ctx.sgemm("R", "N", "N", M, N, K, 1.0, A, 0, B, 0, 0.0, C, 0)
    ref = A.double() @ B.double()
    err = float((C.double() - ref).norm() / ref.norm())
    # all-positive data is the worst case for fp32 accumulation (no cancellation): measured ~1e-6; the tolerance of
    # BASELINE.json is 1e-5
    assert err <= 1e-5, err
    # through the host entry point (row blocks / panels) too
    Ah, Bh = A.cpu().numpy(), B.cpu().numpy()
    Ch = np.full((M, N), np.nan, np.float32)
    ctx.host_gemm("R", "N", "N", M, N, K, 1.0, 0.0, Ah, Bh, Ch)
    assert float(np.linalg.norm(Ch - ref.cpu().numpy()) / np.linalg.norm(ref.cpu().numpy())) <= 1e-5
