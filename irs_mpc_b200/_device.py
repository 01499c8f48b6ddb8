"""Device plumbing: torch is used only for device memory, streams and torch.distributed."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.IrsCudaError(
            "irs_mpc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def to_device(a, dtype=torch.float64):
    """numpy / tensor -> contiguous CUDA tensor of `dtype` (host->device copy if needed)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a), dtype={torch.float64: np.float64,
                                                     torch.float32: np.float32,
                                                     torch.int32: np.int32}[dtype])
    if not arr.flags.writeable:      # e.g. a contiguous slice of np.broadcast_to: torch refuses read-only memory
        arr = arr.copy()
    return torch.from_numpy(arr).to(dev, non_blocking=False)


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def to_numpy(t):
    return t.detach().cpu().numpy()
