"""Parameter containers that mirror the reference's NN blocks (src/models/blocks.py).

These classes own the parameters under exactly the reference's attribute names (so `state_dict()` keys, `Agent.load`
and `utils.configure_opt`'s isinstance-based weight-decay split keep working: every weight lives in an nn.Conv2d /
nn.Linear / nn.GroupNorm).  They do NOT compute anything in PyTorch: the arithmetic of a whole network runs in the
native executor (diamond_b200/csrc), which reads these tensors through their device pointers.  Calling a container
directly raises, so an accidental eager path cannot hide behind the CUDA one.
"""
from typing import List

import torch
from torch import nn

GN_GROUP_SIZE = 32  # blocks.py:12
GN_EPS = 1e-5  # blocks.py:13
ATTN_HEAD_DIM = 8  # blocks.py:14


def conv3x3(cin: int, cout: int) -> nn.Conv2d:  # blocks.py:19
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)


def conv1x1(cin: int, cout: int) -> nn.Conv2d:  # blocks.py:18
    return nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0)


class _NativeOnly(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; it is executed by the native sm_100a executor of the "
            "enclosing model (InnerModel / ActorCritic), not called directly"
        )


class GroupNorm(_NativeOnly):  # blocks.py:24-31
    def __init__(self, in_channels: int) -> None:
        super().__init__()
        self.norm = nn.GroupNorm(max(1, in_channels // GN_GROUP_SIZE), in_channels, eps=GN_EPS)


class AdaGroupNorm(_NativeOnly):  # blocks.py:34-45
    def __init__(self, in_channels: int, cond_channels: int) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.num_groups = max(1, in_channels // GN_GROUP_SIZE)
        self.linear = nn.Linear(cond_channels, in_channels * 2)


class SelfAttention2d(_NativeOnly):  # blocks.py:51-72
    def __init__(self, in_channels: int, head_dim: int = ATTN_HEAD_DIM) -> None:
        super().__init__()
        self.n_head = max(1, in_channels // head_dim)
        assert in_channels % self.n_head == 0
        self.norm = GroupNorm(in_channels)
        self.qkv_proj = conv1x1(in_channels, in_channels * 3)
        self.out_proj = conv1x1(in_channels, in_channels)
        nn.init.zeros_(self.out_proj.weight)
        nn.init.zeros_(self.out_proj.bias)


class FourierFeatures(_NativeOnly):  # blocks.py:78-87
    def __init__(self, cond_channels: int) -> None:
        super().__init__()
        assert cond_channels % 2 == 0
        self.register_buffer("weight", torch.randn(1, cond_channels // 2))


class Downsample(_NativeOnly):  # blocks.py:93-100
    def __init__(self, in_channels: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=1)
        nn.init.orthogonal_(self.conv.weight)


class Upsample(_NativeOnly):  # blocks.py:103-110
    def __init__(self, in_channels: int) -> None:
        super().__init__()
        self.conv = conv3x3(in_channels, in_channels)


class SmallResBlock(_NativeOnly):  # blocks.py:116-123
    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.f = nn.Sequential(GroupNorm(in_channels), nn.SiLU(inplace=True), conv3x3(in_channels, out_channels))
        self.skip_projection = nn.Identity() if in_channels == out_channels else conv1x1(in_channels, out_channels)


class ResBlock(_NativeOnly):  # blocks.py:129-147
    def __init__(self, in_channels: int, out_channels: int, cond_channels: int, attn: bool) -> None:
        super().__init__()
        self.proj = conv1x1(in_channels, out_channels) if in_channels != out_channels else nn.Identity()
        self.norm1 = AdaGroupNorm(in_channels, cond_channels)
        self.conv1 = conv3x3(in_channels, out_channels)
        self.norm2 = AdaGroupNorm(out_channels, cond_channels)
        self.conv2 = conv3x3(out_channels, out_channels)
        self.attn = SelfAttention2d(out_channels) if attn else nn.Identity()
        nn.init.zeros_(self.conv2.weight)


class ResBlocks(_NativeOnly):  # blocks.py:153-177
    def __init__(self, list_in_channels: List[int], list_out_channels: List[int], cond_channels: int, attn: bool) -> None:
        super().__init__()
        assert len(list_in_channels) == len(list_out_channels)
        self.in_channels = list_in_channels[0]
        self.resblocks = nn.ModuleList(
            ResBlock(i, o, cond_channels, attn) for i, o in zip(list_in_channels, list_out_channels)
        )


class UNet(_NativeOnly):  # blocks.py:183-220 (constructor); forward lives in csrc/api.cu PlanBuilder::build
    def __init__(self, cond_channels: int, depths: List[int], channels: List[int], attn_depths: List[int]) -> None:
        super().__init__()
        assert len(depths) == len(channels) == len(attn_depths)
        self._num_down = len(channels) - 1
        d_blocks, u_blocks = [], []
        for i, n in enumerate(depths):
            c1, c2 = channels[max(0, i - 1)], channels[i]
            d_blocks.append(ResBlocks([c1] + [c2] * (n - 1), [c2] * n, cond_channels, attn_depths[i]))
            u_blocks.append(ResBlocks([2 * c2] * n + [c1 + c2], [c2] * n + [c1], cond_channels, attn_depths[i]))
        self.d_blocks = nn.ModuleList(d_blocks)
        self.u_blocks = nn.ModuleList(reversed(u_blocks))
        self.mid_blocks = ResBlocks([channels[-1]] * 2, [channels[-1]] * 2, cond_channels, True)
        self.downsamples = nn.ModuleList([nn.Identity()] + [Downsample(c) for c in channels[:-1]])
        self.upsamples = nn.ModuleList([nn.Identity()] + [Upsample(c) for c in reversed(channels[:-1])])
