"""A few eager training steps of the model /root/reference/main.py builds (for `ncu` launch lists). usage: python tools/mainpy_step.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import trainloop_bench as tb
from src.training.losses import SoftTargetCrossEntropy
from src.training.optim import FusedAdamW
dev = torch.device("cuda:0")
model = tb.build(dev).train()
opt = FusedAdamW(model.parameters(), lr=3e-4, max_grad_norm=1.0)
x = torch.randn(512, 3, 32, 32, device=dev)
la = torch.randint(0, 10, (512,), device=dev)
tgt = torch.nn.functional.one_hot(la, 10).float()
crit = SoftTargetCrossEntropy()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    opt.zero_grad()
    crit(model(x), tgt).backward()
    opt.step()
torch.cuda.synchronize()
print("done")
