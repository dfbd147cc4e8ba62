"""Per-op Python wrappers over the C ABI (NHWC fp32 tensors on a CUDA device).  Thin: they only marshal pointers."""
import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("diamond_b200 ops need contiguous CUDA tensors (no CPU fallback)")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def nchw_to_nhwc(x: Tensor, cpad: Optional[int] = None) -> Tensor:
    _cuda(x)
    b, c, h, w = x.shape
    cp = cpad or c
    out = torch.empty(b, h, w, cp, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().dmd_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), b, c, cp, h * w, _lib.current_stream()))
    return out


def nhwc_to_nchw(x: Tensor, c: Optional[int] = None) -> Tensor:
    _cuda(x)
    b, h, w, cp = x.shape
    c = c or cp
    out = torch.empty(b, c, h, w, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().dmd_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), b, c, cp, h * w, _lib.current_stream()))
    return out


def pack_conv_weight(w: Tensor, cin_pad: int, c0_real: Optional[int] = None, c0_store: Optional[int] = None) -> Tuple[Tensor, int]:
    """torch weight [Cout][Cin][k][k] -> fp16 operand [taps][cin_pad/8][CoutPad][8]; returns (packed, CoutPad)."""
    _cuda(w)
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    cout_pad = round_up(cout, 16)
    c0_real = cin if c0_real is None else c0_real
    c0_store = c0_real if c0_store is None else c0_store
    out = torch.empty(taps * cin_pad * cout_pad, device=w.device, dtype=torch.float16)
    _lib.check(_lib.lib().dmd_pack_conv_weight(w.data_ptr(), out.data_ptr(), cout, cout_pad, cin, cin_pad, taps,
                                              c0_real, c0_store, _lib.current_stream()))
    return out, cout_pad


def gn_stats(x: Tensor, gs: int) -> Tensor:
    """(sum, sumsq) per (image, group) of an NHWC tensor, fp64 [B][C/gs][2]."""
    _cuda(x)
    b, h, w, c = x.shape
    st = torch.zeros(b, c // gs, 2, device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().dmd_gn_stats(x.data_ptr(), st.data_ptr(), b, h * w, c, gs, _lib.current_stream()))
    return st


def conv2d_fprop(src0: Tensor, wpk: Tensor, cout: int, cout_pad: int, cin_pad: int, taps: int = 9, *,
                 src1: Optional[Tensor] = None, bias: Optional[Tensor] = None, upsample: bool = False, stride: int = 1,
                 prologue: int = 0, silu: bool = False, stats0: Optional[Tensor] = None, stats1: Optional[Tensor] = None,
                 gs0: int = 0, gs1: int = 0, film: Optional[Tensor] = None, film_off: int = 0,
                 gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None, eps: float = 1e-5,
                 residual: Optional[Tensor] = None, out_gs: int = 0, debug: int = 0, debug_buf: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    _cuda(src0, src1, wpk, bias, stats0, stats1, film, gamma, beta, residual)
    b, hs, ws, c0 = src0.shape
    h, w = (2 * hs, 2 * ws) if upsample else (hs, ws)
    ho, wo = h // stride, w // stride
    out = torch.empty(b, ho, wo, cout, device=src0.device, dtype=torch.float32)
    ostats = torch.zeros(b, cout // out_gs, 2, device=src0.device, dtype=torch.float64) if out_gs else None
    d = _lib.ConvDesc()
    d.src0, d.src1 = src0.data_ptr(), _lib.ptr(src1)
    d.C0, d.C1, d.Cin = c0, (src1.shape[3] if src1 is not None else 0), cin_pad
    d.B, d.Hs, d.Ws = b, hs, ws
    d.upsample, d.taps, d.stride, d.prologue, d.silu = int(upsample), taps, stride, prologue, int(silu)
    d.stats0, d.stats1, d.gs0, d.gs1 = _lib.ptr(stats0), _lib.ptr(stats1), gs0, gs1
    d.film, d.film_stride, d.film_off = _lib.ptr(film), (film.shape[1] if film is not None else 0), film_off
    d.gamma, d.beta, d.eps = _lib.ptr(gamma), _lib.ptr(beta), eps
    d.wpk, d.bias, d.Cout, d.CoutPad = wpk.data_ptr(), _lib.ptr(bias), cout, cout_pad
    d.residual, d.out, d.out_stats, d.out_gs, d.debug = _lib.ptr(residual), out.data_ptr(), _lib.ptr(ostats), out_gs, debug
    d.debug_buf = _lib.ptr(debug_buf)
    _lib.check(_lib.lib().dmd_conv2d_fprop(C.byref(d), _lib.current_stream()))
    return out, ostats


def attn_fwd(x: Tensor, stats_in: Tensor, gamma: Tensor, beta: Tensor, wqkv: Tensor, bqkv: Tensor, wout: Tensor,
             bout: Tensor, gs: int, eps: float = 1e-5, want_stats: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    _cuda(x, stats_in, gamma, beta, wqkv, bqkv, wout, bout)
    b, h, w, c = x.shape
    out = torch.empty_like(x)
    ostats = torch.zeros(b, c // gs, 2, device=x.device, dtype=torch.float64) if want_stats else None
    _lib.check(_lib.lib().dmd_attn_fwd(x.data_ptr(), stats_in.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                       wqkv.data_ptr(), bqkv.data_ptr(), wout.data_ptr(), bout.data_ptr(), out.data_ptr(),
                                       _lib.ptr(ostats), b, h * w, c, gs, eps, _lib.current_stream()))
    return out, ostats
