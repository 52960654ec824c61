"""K3 parity: tcgen05 GEMM vs torch fp32 matmul of the same bf16 inputs (tolerance: rel-L2 < 6e-3,
max-abs < 2% of the output range — bf16 output rounding is 2^-8 relative)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

SHAPES = [(128, 128, 64), (256, 256, 128), (392, 768, 768), (1000, 2304, 768), (200, 192, 48 + 16), (77, 40, 72),
          (392, 384, 384), (264, 640, 320)]      # the last two: several 128-wide column tiles (ViT-S widths)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_layouts(cuda_device, shape, a_mn, b_mn):
    import gemm_selftest
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand needs a 16-byte aligned leading dimension")
    r = gemm_selftest.run(M, N, K, a_mn, b_mn)
    assert r["ok"], r


@pytest.mark.parametrize("epi", ["bias_relu", "bias_gelu_pre", "bias_res", "relu_mask_res", "gelu_grad", "fp32"])
def test_gemm_epilogues(cuda_device, epi):
    import gemm_selftest
    for (M, N, K) in [(392, 768, 256), (130, 200, 64), (392, 384, 256)]:
        r = gemm_selftest.run(M, N, K, False, False, epi)
        assert r["ok"], r
        if "pre_rel_l2" in r:
            assert r["pre_rel_l2"] < 6e-3


@pytest.mark.parametrize("splits", [2, 5, 0])
def test_gemm_splitk_wgrad_shape(cuda_device, splits):
    import gemm_selftest
    r = gemm_selftest.run(768, 256, 4096, True, True, "fp32", splits)
    assert r["ok"], r
    r = gemm_selftest.run(256, 768, 3000 - 3000 % 8, True, True, "none", splits)
    assert r["ok"], r


@pytest.mark.parametrize("act", ["none", "relu", "gelu"])
def test_gemm_skinny_splitk_bias_activation(cuda_device, act):
    """Weight-streaming shape of the factorised head at small batch (M = 4 rows, long K): split-K with alpha, bias and
    the activation applied by the reduce kernel == the single-pass fused epilogue == torch fp32."""
    import torch
    from sfcvit import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    M, N, K = 4, 512, 16384
    a = (torch.randn(M, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    code = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "gelu": ops.ACT_GELU}[act]
    ref = 0.5 * (a.float() @ w.float().t()) + b.float()
    ref = {"none": lambda t: t, "relu": torch.relu, "gelu": torch.nn.functional.gelu}[act](ref)
    from sfcvit import _lib
    assert _lib.load().sfc_gemm_suggest_splits(M, N, K) > 1
    y_split = ops.gemm(a, w, bias=b, act=code, alpha=0.5, splits=0)
    y_one = ops.gemm(a, w, bias=b, act=code, alpha=0.5, splits=1)
    rel = lambda x: float((x.float() - ref).norm() / ref.norm())
    assert rel(y_split) < 6e-3 and rel(y_one) < 6e-3, (rel(y_split), rel(y_one))


@pytest.mark.parametrize("out_dtype", ["fp32", "bf16"])
def test_gemm_splitk_accumulate(cuda_device, out_dtype):
    """Split-K with accumulate (out += A^T B, the gradient-accumulation form) through the vectorised reduce kernel."""
    import torch
    from sfcvit import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 256, 384, 4096
    a = (torch.randn(K, M, generator=g, device="cuda") * 0.1).to(torch.bfloat16)      # MN-major operands (wgrad shape)
    b = (torch.randn(K, N, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    dt = torch.float32 if out_dtype == "fp32" else torch.bfloat16
    base = torch.randn(M, N, generator=g, device="cuda").to(dt)
    out = base.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=out, splits=4, accumulate=True)
    ref = base.float() + a.float().t() @ b.float()
    assert float((out.float() - ref).norm() / ref.norm()) < (1e-4 if out_dtype == "fp32" else 6e-3)


@pytest.mark.parametrize("shape", [(3072, 768, 50176 // 8), (2304, 768, 3920), (1024, 384, 1568), (256, 768, 136), (384, 96, 1000),
                                   (200, 64, 392), (128, 768, 4096), (1000, 1536, 784)])
@pytest.mark.parametrize("wdt", ["bf16", "fp32"])
def test_wgrad_with_fused_bias_gradient(cuda_device, shape, wdt):
    """ops.wgrad: dW = dY^T X and db = sum_tokens dY. Covered shapes get db as one more accumulator column of the wgrad
    kernel (all-ones B tile; reference: autograd of nn.Linear's bias); the others use the separate column-sum kernel —
    either way the numbers must equal torch fp32 sums of the same bf16 inputs (db: fp32 accumulation, one rounding)."""
    import torch
    from sfcvit import _lib, ops
    N, K, Mtok = shape
    dt = torch.bfloat16 if wdt == "bf16" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(N + K)
    dy = (torch.randn(Mtok, N, generator=g, device="cuda") * 0.5 + 0.05).to(torch.bfloat16)
    x = torch.randn(Mtok, K, generator=g, device="cuda").to(torch.bfloat16)
    fused = _lib.load().sfc_gemm_colsum_splits(N, K, Mtok, 1, 1) > 0
    assert fused == (N > 128 and K % 32 == 0 and Mtok > 64)
    l0 = ops.LAUNCHES
    dw, db = ops.wgrad(dy, x, dt, want_db=True)
    launches = ops.LAUNCHES - l0
    assert launches == (2 if fused else 2 + (2 if _lib.load().sfc_gemm_suggest_splits(N, K, Mtok) > 1 else 1)), launches
    ref_w = dy.float().t() @ x.float()
    ref_b = dy.float().sum(0)
    assert dw.dtype == dt and db.dtype == dt and tuple(dw.shape) == (N, K) and tuple(db.shape) == (N,)
    tol = 6e-3 if wdt == "bf16" else 2e-5
    assert float((dw.float() - ref_w).norm() / ref_w.norm()) < tol
    assert float((db.float() - ref_b).norm() / ref_b.norm()) < (4e-3 if wdt == "bf16" else 1e-5)
    # destinations supplied by the caller (slices of a flat gradient bucket)
    flat = torch.zeros(N * K + N + 64, dtype=dt, device="cuda")
    dw2, db2 = ops.wgrad(dy, x, dt, want_db=True, dw_out=flat[:N * K].view(N, K), db_out=flat[N * K:N * K + N])
    assert dw2.data_ptr() == flat.data_ptr() and torch.equal(dw2, dw) and torch.equal(db2, db)
    assert float(flat[N * K + N:].abs().sum()) == 0.0
