"""Drop-in for torchreid/metrics/distance.py, computed by the sm_100a tensor-core kernel.

Same names, arguments, checks and error behaviour as the reference (distance.py:6-80); the arithmetic
runs in ``libieee_b200.so`` (ieee_b200/csrc/distmat_sm100.cu).
"""
from __future__ import annotations

import torch

from .. import _lib

# fp32 inputs default to the fp32-equivalent f16x3 split; bf16 inputs multiply exactly in one bf16 pass
DEFAULT_PRECISION = {torch.float32: "f16x3", torch.bfloat16: "bf16"}


def _device_distmat(a: torch.Tensor, b: torch.Tensor, metric: str, normalize: bool = False, precision: str | None = None,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """a [Q,D], b [G,D] CUDA tensors (float32 or bfloat16, same dtype) -> float32 [Q,G] on the same device."""
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype and a.dtype in _lib.DTYPES
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    Q, D = a.shape
    G = b.shape[0]
    prec = _lib.PRECISIONS[precision or DEFAULT_PRECISION[a.dtype]]
    if out is None:
        out = torch.empty((Q, G), dtype=torch.float32, device=a.device)
    if Q == 0 or G == 0:
        return out
    lib = _lib.load()
    ws_bytes = lib.ieee_distmat_workspace_bytes(Q, G, D, prec)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("ieee_distmat", a.data_ptr(), b.data_ptr(), _lib.DTYPES[a.dtype], a.stride(0), b.stride(0), Q, G, D,
                  _lib.METRICS.get(metric, _lib.INTERNAL_METRICS.get(metric)), int(normalize), prec, out.data_ptr(),
                  out.stride(0), ws.data_ptr(), ws_bytes,
                  _lib.stream())
    return out


def compute_distance_matrix(input1, input2, metric="euclidean", precision=None):
    """A wrapper function for computing distance matrix (reference: distance.py:6).

    Args:
        input1 (torch.Tensor): 2-D feature matrix.
        input2 (torch.Tensor): 2-D feature matrix.
        metric (str, optional): "euclidean" or "cosine". Default is "euclidean".
        precision (str, optional): "f16x3" (fp32-equivalent, default for float32), "bf16", "fp32_simt".

    Returns:
        torch.Tensor: distance matrix, same dtype and device as the inputs.
    """
    assert isinstance(input1, torch.Tensor)
    assert isinstance(input2, torch.Tensor)
    assert input1.dim() == 2, "Expected 2-D tensor, but got {}-D".format(input1.dim())
    assert input2.dim() == 2, "Expected 2-D tensor, but got {}-D".format(input2.dim())
    assert input1.size(1) == input2.size(1)
    if metric not in ("euclidean", "cosine"):
        raise ValueError(
            'Unknown distance metric: {}. Please choose either "euclidean" or "cosine"'.format(metric))
    _lib.require_cuda()
    src_device, src_dtype = input1.device, input1.dtype
    a, b = input1, input2
    if a.dtype not in _lib.DTYPES:
        a, b = a.float(), b.float()
    if b.dtype != a.dtype:
        b = b.to(a.dtype)
    dev = a.device if a.is_cuda else (b.device if b.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    a = a.to(dev, non_blocking=True)
    b = b.to(dev, non_blocking=True)
    out = _device_distmat(a, b, metric, precision=precision)
    if out.dtype != src_dtype:
        out = out.to(src_dtype)
    return out if src_device.type == "cuda" else out.to(src_device)


def euclidean_squared_distance(input1, input2):
    """Computes euclidean squared distance (reference: distance.py:49-64; squared, unclamped)."""
    return compute_distance_matrix(input1, input2, "euclidean")


def cosine_distance(input1, input2):
    """Computes cosine distance (reference: distance.py:67-80)."""
    return compute_distance_matrix(input1, input2, "cosine")
