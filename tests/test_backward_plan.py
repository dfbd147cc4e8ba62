"""CPU: the formulations planned for the backward kernels (oracle/backward_plan.py) against torch autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle import backward_plan as P


@pytest.mark.parametrize("b,c,co,h,w", [(2, 16, 8, 8, 8), (3, 8, 16, 5, 7), (1, 4, 4, 1, 3)])
def test_wgrad_as_gemm_over_padded_linear_positions(b, c, co, h, w):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, c, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(co, c, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(b, co, h, w, generator=g, dtype=torch.float64)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, gy)
    assert torch.allclose(P.wgrad_over_positions(x, gy), ref, rtol=1e-10, atol=1e-10)


def test_dgrad_as_forward_conv():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 6, 9, 5, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(4, 6, 3, 3, generator=g, dtype=torch.float64)
    gy = torch.randn(2, 4, 9, 5, generator=g, dtype=torch.float64)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), x, gy)
    assert torch.allclose(P.dgrad_as_forward_conv(gy, wt), ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("c", [32, 64, 16])
def test_adagn_silu_backward_two_pass(c):
    g = torch.Generator().manual_seed(3)
    b, h, w = 3, 6, 5
    x = torch.randn(b, c, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    scale = (0.3 * torch.randn(b, c, generator=g, dtype=torch.float64)).requires_grad_(True)
    shift = (0.3 * torch.randn(b, c, generator=g, dtype=torch.float64)).requires_grad_(True)
    gz = torch.randn(b, c, h, w, generator=g, dtype=torch.float64)
    z = F.silu(F.group_norm(x, max(1, c // 32), eps=1e-5) * (1 + scale[:, :, None, None]) + shift[:, :, None, None])
    rx, rs, rt = torch.autograd.grad(z, (x, scale, shift), gz)
    gx, gs, gt = P.adagn_silu_backward_two_pass(x.detach(), scale.detach(), shift.detach(), gz)
    assert torch.allclose(gx, rx, rtol=1e-9, atol=1e-10)
    assert torch.allclose(gs, rs, rtol=1e-9, atol=1e-10)
    assert torch.allclose(gt, rt, rtol=1e-9, atol=1e-10)
