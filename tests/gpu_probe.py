"""Bring-up probe (run on the GPU box, not a pytest test): prints conv errors for the descriptor variants so one
gpurun call is enough to tell a layout problem from an alignment problem."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_conv import _run_conv, _rel

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
for dbg in (0, 1):
    for case in (dict(b=1, h=16, w=16, c0=16, c1=0, cout=16, taps=1), dict(b=1, h=16, w=16, c0=64, c1=0, cout=64, taps=1),
                 dict(b=1, h=16, w=16, c0=16, c1=0, cout=16), dict(b=2, h=64, w=64, c0=64, c1=0, cout=64)):
        try:
            got, ref32, ref16, _ = _run_conv(dev, debug=dbg, **case)
            print(f"dbg={dbg} {case}: rel16={_rel(got, ref16):.3e} rel32={_rel(got, ref32):.3e} finite={bool(torch.isfinite(got).all())}", flush=True)
        except Exception as e:  # noqa
            print(f"dbg={dbg} {case}: EXC {e}", flush=True)
