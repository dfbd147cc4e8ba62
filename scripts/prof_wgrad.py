"""ncu target: the tcgen05 filter-gradient kernel on the dominant backward shape (3x3 64->64 @64x64, `batch` images)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(batch, 64, 64, 64, device=dev)
g = torch.randn(batch, 64, 64, 64, device=dev)
xo, go = ops.prep_act(x)[0], ops.prep_act(g)[0]
dw = torch.zeros(64, 64, 9, device=dev)
for _ in range(4):
    ops.conv2d_wgrad(go, 64, xo, 64, batch, 64, 64, 64, 64, 9, dw=dw)
torch.cuda.synchronize()
print("done")
