"""ncu launch-list target: ONE denoiser training step (Denoiser.forward + backward + clip + AdamW, bench.py's cfg-2 block)
at `batch` samples, no warm-up (ncu serialises and cold-caches every launch anyway)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = bench.train_block(torch.device("cuda:0"), 1, 0, batch, steps=1, warmup=0)
print(out)
