"""The two helpers of the reference's src/utils.py that the mirrored models need (init_lstm utils.py:184-196)."""
from typing import Any, Dict, Tuple

import torch.nn as nn
from torch import Tensor

LossAndLogs = Tuple[Tensor, Dict[str, Any]]


def init_lstm(model: nn.Module) -> None:
    for name, p in model.named_parameters():
        if "weight_ih" in name:
            nn.init.xavier_uniform_(p.data)
        elif "weight_hh" in name:
            nn.init.orthogonal_(p.data)
        elif "bias_ih" in name:
            p.data.fill_(0)
            n = p.size(0)
            p.data[(n // 4):(n // 2)].fill_(1)  # forget-gate bias
        elif "bias_hh" in name:
            p.data.fill_(0)


class NativeStateMixin:
    """Modules with a native executor check on every call whether a parameter changed (data_ptr, _version) before reusing the
    packed fp16 copies.  Walking state_dict() for that costs ~0.4 ms per call on the 235 tensors of the denoiser; the tensor
    list is therefore cached and dropped whenever `_apply` (to / cuda / float ...) may have replaced tensors.  In-place updates
    (optimizer steps, load_state_dict, p.data.copy_) keep the objects and bump `_version`, which the key sees; code that REPLACES
    a Parameter object or writes through `.data.fill_` must call `refresh_weights()`."""

    def _state_tensors(self):
        ts = self.__dict__.get("_state_tensor_cache")
        if ts is None:
            ts = list(self.state_dict(keep_vars=True).values())
            self.__dict__["_state_tensor_cache"] = ts
        return ts

    def refresh_weights(self) -> None:
        self.__dict__["_state_tensor_cache"] = None
        self.__dict__["_wkey"] = None

    def _apply(self, fn, recurse=True):
        self.__dict__["_state_tensor_cache"] = None
        return super()._apply(fn, recurse)

    # The native handle (`_h`, a raw pointer owned by __del__), the packed fp16 weights, workspaces and cached layouts belong
    # to ONE module object.  copy.deepcopy (EMA copies) and pickle (multiprocessing) go through __getstate__: the copy starts
    # without native state and builds its own on first use, so two objects never own -- and free -- the same handle.
    _NATIVE_RESET = ("_h", "_h_key", "_h_dev", "_wkey", "_packed", "_ws")
    _NATIVE_DROP = ("_state_tensor_cache", "_tws_pool", "_ws_pool", "_bwd_scratch", "_grad_acc", "_gv_layout", "last_flat_grad")

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in self._NATIVE_RESET:
            if k in state:
                state[k] = None
        for k in self._NATIVE_DROP:
            state.pop(k, None)
        return state

    @staticmethod
    def require_current_device(dev) -> None:
        """The native layer launches on the CURRENT CUDA device and on its current stream; a model that lives on another device
        would silently run on the wrong GPU.  Fail loudly instead (the reference trainer calls torch.cuda.set_device, trainer.py:52)."""
        import torch

        if dev.index is not None and dev.index != torch.cuda.current_device():
            raise RuntimeError(f"diamond_b200: the module is on {dev} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               f"call torch.cuda.set_device({dev.index}) (or wrap the call in torch.cuda.device) first")


# ----------------------------------------------------------------------------------------------- data-parallel plumbing
# SURVEY.md 8 a26.  The reference wraps each agent module in torch DDP (utils.py:105-106, trainer.py:110) and relies on
# autograd hooks to average gradients.  A native executor produces all gradients of a module in one C-ABI call, outside
# autograd, so the averaging is explicit: flat fp32 buckets, one all_reduce per bucket (NCCL over NVLink on GPUs, gloo in
# the CPU tests), same result as DDP (sum over ranks / world size).


def broadcast_if_needed(*args):
    """utils.py:97-102: every rank ends up with rank 0's objects; a no-op without a process group."""
    import torch.distributed as dist

    objects = list(args)
    if dist.is_available() and dist.is_initialized():
        dist.broadcast_object_list(objects, src=0)
    return objects


def allreduce_gradients(params, bucket_bytes: int = 32 << 20, group=None) -> int:
    """Average `.grad` of `params` over the process group in flat buckets of <= bucket_bytes (one collective each; with
    NVSwitch the cost is launch latency, not link count, so buckets are large).  Parameters without a gradient are treated
    as zero on this rank (DDP's find_unused_parameters semantics) so that every rank issues identical collectives.
    Returns the number of all_reduce calls issued (0 without a process group)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    calls, i = 0, 0
    while i < len(params):
        j, nbytes = i, 0
        while j < len(params) and (j == i or nbytes + params[j].numel() * 4 <= bucket_bytes):
            nbytes += params[j].numel() * 4
            j += 1
        chunk = params[i:j]
        flat = torch.zeros(sum(p.numel() for p in chunk), dtype=torch.float32, device=chunk[0].device)
        off = 0
        for p in chunk:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for p in chunk:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        calls += 1
        i = j
    return calls


def allreduce_native_gradients(inner_module, group=None) -> int:
    """Data-parallel gradient averaging for a model whose backward ran natively (diamond_b200 InnerModel): the native
    backward wrote EVERY parameter gradient into one flat fp32 buffer and autograd adopted views of it as `.grad`, so the whole
    model is averaged by ONE all_reduce on that buffer (NCCL over NVLink / NVSwitch; the reference wraps each model in DDP,
    utils.py:105-106, which buckets the same bytes into several collectives).  Falls back to `allreduce_gradients` (flatten,
    reduce, scatter) when the gradients do not alias the flat buffer (e.g. after gradient accumulation).  Returns the number of
    collectives issued."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = getattr(inner_module, "last_flat_grad", None)
    params = [p for p in inner_module.parameters() if p.requires_grad]
    aliased = flat is not None and all(
        p.grad is not None and flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in params)
    if not aliased:
        return allreduce_gradients(params, group=group)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return 1
