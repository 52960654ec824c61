"""Stand-alone GEMM check (one variant per process so a faulting kernel cannot poison later checks).
usage: python tools/gemm_selftest.py M N K a_mn b_mn [epi] [splits]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import torch  # noqa: E402
from sfcvit import ops  # noqa: E402


def run(M, N, K, a_mn, b_mn, epi="none", splits=1, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    A = (torch.randn(M, K, generator=g, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, generator=g, device=dev) * 0.5).bfloat16()
    ref = A.float() @ B.float().t()
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    kw = {}
    if epi == "bias_relu":
        bias = torch.randn(N, generator=g, device=dev).bfloat16()
        kw = dict(bias=bias, act=ops.ACT_RELU)
        ref = torch.relu(ref + bias.float())
    elif epi == "bias_gelu_pre":
        bias = torch.randn(N, generator=g, device=dev).bfloat16()
        kw = dict(bias=bias, act=ops.ACT_GELU, want_pre=True)
        pre_ref = ref + bias.float()
        ref = torch.nn.functional.gelu(pre_ref)
    elif epi == "bias_res":
        bias = torch.randn(N, generator=g, device=dev).bfloat16()
        res = torch.randn(M, N, generator=g, device=dev).bfloat16()
        kw = dict(bias=bias, residual=res)
        ref = ref + bias.float() + res.float()
    elif epi == "relu_mask_res":
        aux = torch.randn(M, N, generator=g, device=dev).bfloat16()
        res = torch.randn(M, N, generator=g, device=dev).bfloat16()
        kw = dict(aux=aux, aux_mode=ops.AUX_RELU_MASK, residual=res)
        ref = ref * (aux.float() > 0) + res.float()
    elif epi == "gelu_grad":
        aux = torch.randn(M, N, generator=g, device=dev).bfloat16()
        kw = dict(aux=aux, aux_mode=ops.AUX_GELU_GRAD)
        x = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        ref = ref * x.grad
    elif epi == "fp32":
        kw = dict(out_dtype=torch.float32)
    out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, splits=splits, **kw)
    pre = None
    if isinstance(out, tuple):
        out, pre = out
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm().clamp_min(1e-9)).item()
    res = dict(M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, epi=epi, splits=splits, max_abs_err=err, ref_max=scale, rel_l2=rel)
    if pre is not None:
        res["pre_rel_l2"] = ((pre.float() - pre_ref).norm() / pre_ref.norm()).item()
    res["ok"] = bool(rel < 6e-3 and err <= 2e-2 * max(scale, 1.0))
    return res


if __name__ == "__main__":
    a = sys.argv[1:]
    M, N, K, a_mn, b_mn = int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4])
    epi = a[5] if len(a) > 5 else "none"
    splits = int(a[6]) if len(a) > 6 else 1
    print(json.dumps(run(M, N, K, bool(a_mn), bool(b_mn), epi, splits)))
