"""GPU parity: tcgen05 GEMM vs a torch fp32 reference on the same 16-bit operands."""
import numpy as np
import pytest

from spittle_b200 import capi

pytestmark = pytest.mark.gpu


def _run(dev, dtype, M, N, K, out_f32, bias, act, residual, res_row_mod=0, seed=0):
    import torch
    tdt = torch.bfloat16 if dtype == capi.SB_DTYPE_BF16 else torch.float16
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).to(tdt).to(dev)
    W = (torch.randn(N, K, generator=g) * 0.1).to(tdt).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    rrows = res_row_mod if res_row_mod else M
    R = torch.randn(rrows, N, generator=g).to(dev) if residual else None
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else tdt, device=dev)
    capi.gemm_tn_dev(dtype, A.data_ptr(), K, W.data_ptr(), K, M, N, K, out.data_ptr(), N, out_f32,
                     b.data_ptr() if bias else 0, act, R.data_ptr() if residual else 0, N, res_row_mod,
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T
    if bias:
        ref = ref + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if residual:
        idx = torch.arange(M, device=dev) % rrows
        ref = ref + R[idx]
    return out.float(), ref


@pytest.mark.parametrize("dtype", [capi.SB_DTYPE_BF16, capi.SB_DTYPE_F16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 768), (300, 768, 768), (1500, 2304, 768),
                                   (1500 * 3, 3072, 768), (257, 128, 128), (1000, 776, 240), (64, 40, 72)])
def test_gemm_plain(cuda_dev, dtype, M, N, K):
    out, ref = _run(cuda_dev, dtype, M, N, K, True, False, 0, False)
    # fp32 accumulate of exact 16-bit products: only summation order differs
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    assert not out.isnan().any()


@pytest.mark.parametrize("dtype", [capi.SB_DTYPE_BF16, capi.SB_DTYPE_F16])
def test_gemm_epilogues(cuda_dev, dtype):
    import torch
    eps16 = 2 ** -8 if dtype == capi.SB_DTYPE_BF16 else 2 ** -11
    # bias + GELU, 16-bit out
    out, ref = _run(cuda_dev, dtype, 1500, 3072, 768, False, True, 1, False, seed=1)
    assert ((out - ref).abs() <= eps16 * ref.abs() + 2e-3).all()
    # bias + residual, f32 out
    out, ref = _run(cuda_dev, dtype, 1500, 768, 3072, True, True, 0, True, seed=2)
    assert (out - ref).abs().max().item() <= 5e-3
    # bias + GELU + positional residual (row % 1500), f32 out
    out, ref = _run(cuda_dev, dtype, 3000, 768, 2304, True, True, 1, True, res_row_mod=1500, seed=3)
    assert (out - ref).abs().max().item() <= 5e-3


def test_gemm_inplace_residual_and_many_tiles(cuda_dev):
    """x += A W^T + b with out aliasing the residual, more tiles than SMs (persistent loop)."""
    import torch
    dev = cuda_dev
    M, N, K = 1500 * 16, 768, 768
    g = torch.Generator(device="cpu").manual_seed(5)
    A = (torch.randn(M, K, generator=g) * 0.5).bfloat16().to(dev)
    W = (torch.randn(N, K, generator=g) * 0.05).bfloat16().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    ref = x + A.float() @ W.float().T + b
    capi.gemm_tn_dev(capi.SB_DTYPE_BF16, A.data_ptr(), K, W.data_ptr(), K, M, N, K, x.data_ptr(), N, True,
                     b.data_ptr(), 0, x.data_ptr(), N, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() <= 5e-3
