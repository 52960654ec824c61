"""Per-kernel timing at the ViT-B/16 224 px training shapes (CUDA events, warm, rotating operands larger than L2).
usage: python tools/op_bench.py [B] [what ...]     what in {gemm, attn, ln, misc, patch}; default all
Prints one JSON line per case: ms, TFLOP/s or GB/s and the fraction of MEASURED_PEAKS.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sfcvit import ops  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
DEV = "cuda"


def timeit(fn, nrot, iters=12, warm=3):
    for i in range(warm):
        fn(i % nrot)
    torch.cuda.synchronize()
    if os.environ.get("OPBENCH_GRAPH"):           # replay a captured rotation: no host launch cost in the number
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nrot):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (4 * nrot)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rnd(*shape, scale=0.5):
    return (torch.randn(*shape, device=DEV) * scale).bfloat16()


def report(name, ms, flops=None, nbytes=None, **extra):
    r = dict(case=name, ms=round(ms, 4))
    if flops:
        r["tflops"] = round(flops / ms / 1e9, 1)
        r["frac_tensor"] = round(flops / ms / 1e9 / PEAKS["bf16_tflops_sustained"], 3)
    if nbytes:
        r["gbs"] = round(nbytes / ms / 1e6, 1)
        r["frac_hbm"] = round(nbytes / ms / 1e6 / PEAKS["hbm_gbs"], 3)
    r.update(extra)
    print(json.dumps(r), flush=True)


def bench_gemm(M, D=768, F=3072):
    nrot = 3
    shapes = [("qkv", 3 * D, D), ("out_proj", D, D), ("ff1", F, D), ("ff2", D, F)]
    for name, N, K in shapes:
        xs = [rnd(M, K) for _ in range(nrot)]
        w = rnd(N, K, scale=0.05)
        bias = rnd(N)
        res = rnd(M, N)
        dys = [rnd(M, N) for _ in range(nrot)]
        fl = 2.0 * M * N * K
        if os.environ.get("OPBENCH_CUBLAS"):
            report(f"cublas_{name}_plain", timeit(lambda i: torch.matmul(xs[i], w.t()), nrot), fl, M=M, N=N, K=K)
            report(f"cublas_{name}_bias", timeit(lambda i: torch.nn.functional.linear(xs[i], w, bias), nrot), fl)
            report(f"cublas_{name}_dgrad", timeit(lambda i: torch.matmul(dys[i], w), nrot), fl)
            report(f"cublas_{name}_wgrad", timeit(lambda i: torch.matmul(dys[i].t(), xs[i]), nrot), fl)
        report(f"fwd_{name}_plain", timeit(lambda i: ops.gemm(xs[i], w), nrot), fl, M=M, N=N, K=K)
        report(f"fwd_{name}_bias", timeit(lambda i: ops.gemm(xs[i], w, bias=bias), nrot), fl)
        if name == "ff1":
            report(f"fwd_{name}_bias_relu_drop", timeit(lambda i: ops.gemm(xs[i], w, bias=bias, act=ops.ACT_RELU, drop_p=0.1, drop_seed=7), nrot), fl)
        if name in ("out_proj", "ff2"):
            report(f"fwd_{name}_bias_res_drop", timeit(lambda i: ops.gemm(xs[i], w, bias=bias, residual=res, drop_p=0.1, drop_seed=7), nrot), fl)
        report(f"dgrad_{name}", timeit(lambda i: ops.gemm(dys[i], w, b_mn=True), nrot), fl)
        if name == "ff2":
            aux = rnd(M, K)
            report(f"dgrad_{name}_relumask", timeit(lambda i: ops.gemm(dys[i], w, b_mn=True, aux=aux, aux_mode=ops.AUX_RELU_MASK, alpha=1.1), nrot), fl)
        for sp in (0, 1):
            report(f"wgrad_{name}_splits{sp}", timeit(lambda i: ops.gemm(dys[i], xs[i], a_mn=True, b_mn=True, splits=sp), nrot), fl,
                   splits=ops._lib.load().sfc_gemm_suggest_splits(N, K, M) if sp == 0 else 1)
        report(f"wgrad_{name}_with_db_fused", timeit(lambda i: ops.wgrad(dys[i], xs[i], torch.bfloat16, want_db=True), nrot), fl)
        report(f"colsum_{name}_separate", timeit(lambda i: ops.colsum(dys[i]), nrot), None, M * N * 2)
        del xs, dys, res


def bench_attn(B, H=12, N=196):
    D = H * 64
    nrot = 3
    qkvs = [rnd(B * N, 3 * D, scale=1.0) for _ in range(nrot)]
    for dp in (0.0, 0.1):
        fl = 4.0 * N * N * D * B
        by = B * N * D * 2 * 4
        report(f"attn_fwd_drop{dp}", timeit(lambda i: ops.attn_fwd(qkvs[i], B, H, N, drop_p=dp, drop_seed=5), nrot), fl, by, N=N)
        out, lse = ops.attn_fwd(qkvs[0], B, H, N, drop_p=dp, drop_seed=5)
        dout = rnd(B * N, D)
        report(f"attn_bwd_drop{dp}", timeit(lambda i: ops.attn_bwd(qkvs[0], out, dout, lse, B, H, N, drop_p=dp, drop_seed=5), nrot),
               2.5 * fl, B * N * D * 2 * 8)


def bench_ln(M, D=768):
    nrot = 4
    xs = [rnd(M, D) for _ in range(nrot)]
    dys = [rnd(M, D) for _ in range(nrot)]
    g, b = rnd(D), rnd(D)
    report("ln_fwd", timeit(lambda i: ops.layernorm_fwd(xs[i], g, b), nrot), None, M * D * 2 * 2)
    y, mean, rstd = ops.layernorm_fwd(xs[0], g, b)
    report("ln_bwd_plain", timeit(lambda i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, g), nrot), None, M * D * 2 * 3)
    report("ln_bwd_drop_csum", timeit(lambda i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, g, drop_p=0.1, drop_seed=3, want_colsum=True), nrot),
           None, M * D * 2 * 4)


def bench_misc(M, D=768, F=3072):
    nrot = 3
    for N in (D, 3 * D, F):
        xs = [rnd(M, N) for _ in range(nrot)]
        report(f"colsum_{N}", timeit(lambda i: ops.colsum(xs[i]), nrot), None, M * N * 2)
        del xs
    n = 86_000_000
    p, g_ = rnd(n), rnd(n)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    st = torch.zeros(1, device=DEV)
    report("adamw_86M_bf16", timeit(lambda i: ops.adamw_step(p, g_, m, v, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=i + 1,
                                                             max_norm=1.0, stats=st), 1), None, n * 2 * 7)
    report("sumsq_86M_bf16", timeit(lambda i: ops.grad_sumsq(g_, st), 1), None, n * 2)


def bench_patch(B, D=768):
    from sfcvit import functional as F_
    nrot = 3
    for dt in (torch.float32, torch.bfloat16):
        imgs = [torch.randn(B, 3, 224, 224, device=DEV).to(dt) for _ in range(nrot)]
        perm, _ = ops.curve_perm("hilbert", 14, 14, DEV)
        w = torch.randn(D, 768, device=DEV) * 0.02
        wk = F_.kernel_weight(w, 3, 16, 1, "p1p2c")
        bias = rnd(D)
        by = B * (3 * 224 * 224 * imgs[0].element_size() + 196 * D * 2)
        report(f"patch_embed_fwd_{str(dt)[6:]}", timeit(lambda i: ops.patch_embed_fwd(imgs[i], perm, wk, bias, 16, 1), nrot), 2.0 * B * 196 * 768 * D, by)
        report(f"patch_gather_{str(dt)[6:]}", timeit(lambda i: ops.patch_gather(imgs[i], perm, 16, 1), nrot), None,
               B * (3 * 224 * 224 * imgs[0].element_size() + 196 * 768 * 2))
        del imgs


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    what = sys.argv[2:] or ["gemm", "attn", "ln", "misc", "patch"]
    M = B * 196
    if "gemm" in what:
        bench_gemm(M)
    if "attn" in what:
        bench_attn(B)
    if "ln" in what:
        bench_ln(M)
    if "misc" in what:
        bench_misc(M)
    if "patch" in what:
        bench_patch(B)
