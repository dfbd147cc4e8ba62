"""Bring-up diagnostic: the conv kernel's epilogue features one at a time, each case in its own process with a watchdog.
usage: python scripts/diag_conv.py            (driver)   |   python scripts/diag_conv.py CASE   (worker)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "plain64": dict(shape=dict(b=2, h=64, w=64, c0=64, c1=0, cout=64)),
    "stats64": dict(shape=dict(b=2, h=64, w=64, c0=64, c1=0, cout=64), want_stats=True),
    "resid64": dict(shape=dict(b=2, h=64, w=64, c0=64, c1=0, cout=64), residual=True),
    "prol64": dict(shape=dict(b=2, h=64, w=64, c0=64, c1=0, cout=64), prologue=1, silu=True),
    "all64": dict(shape=dict(b=2, h=64, w=64, c0=64, c1=0, cout=64), prologue=1, silu=True, residual=True, want_stats=True),
    "stats16cat": dict(shape=dict(b=3, h=16, w=16, c0=64, c1=64, cout=64), want_stats=True),
    "all16cat": dict(shape=dict(b=3, h=16, w=16, c0=64, c1=64, cout=64), prologue=1, silu=True, residual=True, want_stats=True),
    "all8": dict(shape=dict(b=4, h=8, w=8, c0=64, c1=0, cout=64), prologue=1, silu=True, residual=True, want_stats=True),
}


def worker(name):
    import ctypes

    import torch
    import test_gpu_conv as T

    if os.environ.get("DMD_LIB"):   # control: run the same case on another build of the library
        from diamond_b200 import _lib
        _lib.LIB_PATH = os.environ["DMD_LIB"]
        handle = ctypes.CDLL(_lib.LIB_PATH)
        _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(handle, k) and k != "dmd_sampler_sample"}

    c = CASES[name]
    kw = {k: v for k, v in c.items() if k != "shape"}
    got, ref32, ref16, st = T._run_conv(torch.device("cuda:0"), seed=3, **c["shape"], **kw)
    msg = f"rel32={T._rel(got, ref32):.2e}"
    if st is not None:
        b, ch, h, w = got.shape
        v = got.double().reshape(b, ch // 32, 32 * h * w)
        want = torch.stack([v.sum(-1), (v * v).sum(-1)], -1)
        msg += f" stats_err={float((st - want).abs().max() / want.abs().max()):.2e}"
    if kw.get("prologue"):   # the prologue's inputs: GroupNorm partial sums of the source
        from diamond_b200 import ops
        g = torch.Generator().manual_seed(3)
        sh = c["shape"]
        x0 = torch.randn(sh["b"], sh["c0"], sh["h"], sh["w"], generator=g)
        s0 = ops.nchw_to_nhwc(x0.cuda())
        stx = ops.gn_stats(s0, 32).cpu()
        v = x0.double().reshape(sh["b"], sh["c0"] // 32, -1)
        want = torch.stack([v.sum(-1), (v * v).sum(-1)], -1)
        msg += f" gn_stats_err={float((stx - want).abs().max() / want.abs().max()):.2e}"
    print(name, os.environ.get("DMD_CONV_GROUPS"), msg, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        worker(sys.argv[1])
        sys.exit(0)
    runs = [("1", None), ("2", None)]
    ctl = os.path.join(ROOT, "diamond_b200", "_r1_control.so")
    if os.path.exists(ctl):
        runs.insert(0, ("r1-control", ctl))
    only = os.environ.get("DIAG_CASES", "").split(",") if os.environ.get("DIAG_CASES") else list(CASES)
    for groups, libpath in runs:
        for name in only:
            env = dict(os.environ, DMD_CONV_GROUPS=groups if libpath is None else "1")
            if libpath:
                env["DMD_LIB"] = libpath
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], env=env, capture_output=True, text=True, timeout=60)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                err = (r.stderr.strip().splitlines() or [""])[-1] if r.returncode else ""
                print(f"groups={groups} {name}: rc={r.returncode} {tail} {err[:200]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"groups={groups} {name}: TIMEOUT (hang)", flush=True)
