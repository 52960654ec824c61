"""Throughput of the reference's OWN entry path (/root/reference/main.py:269-314) on the B200 kernels: the model main.py
builds (HierarchicalMortonEmbedding(32, 3, [16, 4, 1], 256) -> VisionTransformer1D depth 8, 4 heads, mlp 512, 10 classes,
default dtype bf16, torch.compile(mode="reduce-overhead") wrapper), batch 512, torch.optim.AdamW + cosine schedule, driven
through src.training.train.train_with_mixup_or_cutmix — next to the same model under GraphedStep + FusedAdamW (the path
bench.py times on ViT-B). usage: python tools/trainloop_bench.py [steps]   -> one JSON line"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def build(device):
    from src.models.vit import VisionTransformer1D
    from src.tokenizers.multiscale.multi_morton import HierarchicalMortonEmbedding
    torch.manual_seed(42)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        pe = HierarchicalMortonEmbedding(img_size=32, in_channels=3, patch_size_list=[16, 4, 1], embed_dim=256)
        model = VisionTransformer1D(patch_embed=pe, depth=8, n_heads=4, mlp_dim=512, num_classes=10).to(device)
    finally:
        torch.set_default_dtype(prev)
    return model


class Loader(list):
    dataset = None


def epoch_through_train_py(device, steps, B, graph):
    from src.training.losses import SoftTargetCrossEntropy
    from src.training.train import train_with_mixup_or_cutmix
    os.environ["SFC_TRAIN_GRAPH"] = "1" if graph else "0"
    model = torch.compile(build(device), mode="reduce-overhead")      # main.py:284
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.00005)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    g = torch.Generator().manual_seed(0)
    data = [(torch.randn(B, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 10, (B,), generator=g).pin_memory()) for _ in range(4)]
    np.random.seed(0)
    crit = SoftTargetCrossEntropy()
    warm = Loader([data[i % 4] for i in range(4)]); warm.dataset = list(range(4 * B))
    train_with_mixup_or_cutmix(model, warm, crit, opt, sched, device)
    run = Loader([data[i % 4] for i in range(steps)]); run.dataset = list(range(steps * B))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss, acc = train_with_mixup_or_cutmix(model, run, crit, opt, sched, device)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return steps * B / dt, dt / steps * 1e3, loss


def graphed_fused(device, steps, B):
    from src.training.graphs import GraphedStep
    from src.training.losses import SoftTargetCrossEntropy
    from src.training.optim import FusedAdamW
    model = build(device).train()
    opt = FusedAdamW(model.parameters(), lr=3e-4, weight_decay=0.00005, max_grad_norm=1.0)
    g = torch.Generator(device=device).manual_seed(0)
    xs = [torch.randn(B, 3, 32, 32, generator=g, device=device) for _ in range(4)]
    la = torch.randint(0, 10, (B,), generator=g, device=device)
    tgt = 0.3 * torch.nn.functional.one_hot(la, 10).float() + 0.7 * torch.nn.functional.one_hot(la.roll(1), 10).float()
    step = GraphedStep(model, SoftTargetCrossEntropy(), xs[0], tgt, optimizer=opt)
    for i in range(4):
        step(xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return B / (ms * 1e-3), ms


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda:0")
    B = 512
    a_ips, a_ms, a_loss = epoch_through_train_py(dev, steps, B, graph=True)
    b_ips, b_ms, _ = epoch_through_train_py(dev, steps, B, graph=False)
    c_ips, c_ms = graphed_fused(dev, steps, B)
    print(json.dumps({"workload": "main.py model (hier-Morton 32px, 3x256, depth 8, 4 heads, B 512), one epoch slice", "steps": steps,
                      "train_with_mixup_or_cutmix (lazy CUDA graph, torch AdamW, H2D + .item() per step)": {"images_per_s": a_ips, "ms_per_step": a_ms, "loss": a_loss},
                      "same loop, eager launches (SFC_TRAIN_GRAPH=0)": {"images_per_s": b_ips, "ms_per_step": b_ms},
                      "GraphedStep + FusedAdamW (device-resident inputs)": {"images_per_s": c_ips, "ms_per_step": c_ms},
                      "entry_path_vs_graphed_fused": a_ips / c_ips}))
