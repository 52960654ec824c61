"""On-chip rate micro-benchmarks (TMEM load bandwidth, MUFU.EX2 rate). usage: python tools/debug_bench.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200"))
import subprocess
import torch
# built on demand into tools/ (NOT part of the product library libsfcvit.so)
PK = os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200")
so = os.path.join(ROOT, "tools", "libsfcdebug.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
                       "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PK, "csrc"),
                       os.path.join(ROOT, "tools", "debug_bench.cu"), os.path.join(PK, "csrc", "host.cu"), "-lcuda", "-o", so])
lib = ctypes.CDLL(so)
f = lib.sfc_debug_bench
f.restype = ctypes.c_int
cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
sink = torch.zeros(1, device="cuda")
vp = ctypes.c_void_p
for which, name in ((0, "tmem_ld"), (1, "ex2")):
    for threads in (128, 256, 512):
        iters = 200
        for _ in range(2):
            rc = f(which, threads, iters, vp(cyc.data_ptr()), vp(sink.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
            assert rc == 0
            torch.cuda.synchronize()
        c = float(cyc.float().mean())
        if which == 0:
            nbytes = (threads // 32) * 32 * 512 * 4 * iters
            print(f"{name} threads={threads}: {c:.0f} cycles, {nbytes / c:.1f} B/clk/SM")
        else:
            n = threads * 64 * iters
            print(f"{name} threads={threads}: {c:.0f} cycles, {n / c:.2f} ex2/clk/SM")
