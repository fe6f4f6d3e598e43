"""Batched grid shortest paths on the GPU — the replacement for the reference's ``spfa.spfa``
(deps/spfa/src/main.cpp:10-116; pybind signature ``spfa(input_map: bool[H,W], source: (i, j)) ->
(dists f32[H,W], parents i32[H,W])``)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def spfa_batch(maps: torch.Tensor, sources: torch.Tensor, reference_parents: bool = False):
    """maps: [B,H,W] (any dtype; non-zero = free, like pybind's bool cast), sources: [B,2] (row, col).
    Returns (dists f32 [B,H,W], parents i32 [B,H,W]) on the same CUDA device, enqueued on the current stream.
    reference_parents: parents in the reference solver's own tie-break order (bp_spfa_batch_exact: one thread per map restates
    its queue; slow, synchronises); the default parents are a valid shortest-path tree with a fixed direction-order tie-break."""
    if not maps.is_cuda:
        raise _lib.BenchPushError("spfa_batch needs CUDA tensors (there is no CPU fallback)")
    B, H, W = maps.shape
    m = (maps != 0).to(torch.uint8).contiguous()
    s = sources.to(device=maps.device, dtype=torch.int32).contiguous()
    if int(s[:, 0].min()) < 0 or int(s[:, 0].max()) >= H or int(s[:, 1].min()) < 0 or int(s[:, 1].max()) >= W:
        raise ValueError("source out of bounds")
    d = torch.empty((B, H, W), dtype=torch.float32, device=maps.device)
    p = torch.empty((B, H, W), dtype=torch.int32, device=maps.device)
    with torch.cuda.device(maps.device):
        st = C.c_void_p(torch.cuda.current_stream(maps.device).cuda_stream)
        fn = _lib.lib().bp_spfa_batch_exact if reference_parents else _lib.lib().bp_spfa_batch
        _lib.check(fn(C.c_void_p(m.data_ptr()), C.c_void_p(s.data_ptr()), B, H, W,
                                            C.c_void_p(d.data_ptr()), C.c_void_p(p.data_ptr()), st))
    return d, p


def spfa(input_map, source, reference_parents: bool = False):
    """Drop-in for ``spfa.spfa(input_map, source)`` (numpy in, numpy out) running on the GPU."""
    m = np.ascontiguousarray(np.asarray(input_map).astype(bool).astype(np.uint8))
    H, W = m.shape
    if not (0 <= int(source[0]) < H and 0 <= int(source[1]) < W):
        raise ValueError("source out of bounds")
    src = np.array([[int(source[0]), int(source[1])]], dtype=np.int32)
    d = np.empty((H, W), dtype=np.float32)
    p = np.empty((H, W), dtype=np.int32)
    fn = _lib.lib().bp_spfa_batch_exact_host if reference_parents else _lib.lib().bp_spfa_batch_host
    _lib.check(fn(m.ctypes.data, src.ctypes.data, 1, H, W, d.ctypes.data, p.ctypes.data))
    return d, p
