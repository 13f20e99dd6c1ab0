"""Multi-GPU host logic: depth frames are independent units, so the path shards
with NO per-frame collective.  One process per GPU (torch.distributed: nccl on GPUs, gloo in the
CPU tests); the only communication is one broadcast per static buffer at set-up (mesh + initial
pose/frames) and, optionally, an ordered gather of per-frame results on the host.

torch is plumbing here (process group + device tensors); no kernel of the path lives in this file.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> list[int]:
    """C5 (one stream, N GPUs): frame k -> GPU k mod N."""
    return list(range(rank, n_frames, world))


def streams_for_rank(n_streams: int, rank: int, world: int) -> list[int]:
    """C4 (N camera streams): stream s is pinned to GPU s mod N."""
    return [s for s in range(n_streams) if s % world == rank]


def _bcast(t: torch.Tensor, src: int):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src)
    return t


def broadcast_arrays(arrays: dict | None, src: int, device, rank: int | None = None) -> dict:
    """Broadcast a dict of numpy arrays from `src` to every rank as device tensors: a header
    (names are fixed by the caller's order; shapes/dtypes travel in an int64 tensor), then one
    broadcast per buffer.  On `src`, `arrays` holds the data; elsewhere it may be None."""
    rank = dist.get_rank() if rank is None and dist.is_initialized() else (rank or 0)
    # uint16 (16UC1 depth) and uint32 (tri_part of ruf_set_model) travel as their signed twins: torch has no
    # arithmetic on them, and the payload is moved as raw bytes anyway
    dtypes = [np.float32, np.float64, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32]
    as_torch = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32}
    MAXN, MAXD = 16, 6
    # row MAXN carries the entry count explicitly (a 0-d array is a legal entry: ndim alone cannot say "absent")
    hdr = torch.zeros((MAXN + 1, MAXD + 2), dtype=torch.int64, device=device)
    names = None
    if rank == src:
        names = list(arrays.keys())
        assert len(names) <= MAXN
        for i, k in enumerate(names):
            a = arrays[k] = np.asarray(arrays[k])
            assert a.ndim <= MAXD, f"{k}: more than {MAXD} dimensions"
            hdr[i, 0] = a.ndim
            hdr[i, 1] = [np.dtype(d) for d in dtypes].index(a.dtype)
            for j, s in enumerate(a.shape):
                hdr[i, 2 + j] = s
        hdr[MAXN, 0] = len(names)
    _bcast(hdr, src)
    n = int(hdr[MAXN, 0])
    obj = [names]
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast_object_list(obj, src, device=torch.device(device) if not isinstance(device, torch.device) else device)
    names = obj[0]
    out = {}
    h = hdr.cpu().numpy()
    for i in range(n):
        nd, dt = int(h[i, 0]), np.dtype(dtypes[int(h[i, 1])])
        dt = np.dtype(as_torch.get(dt, dt))
        shape = tuple(int(v) for v in h[i, 2:2 + nd])
        if rank == src:
            a = arrays[names[i]]
            a = a if a.ndim == 0 else np.ascontiguousarray(a)      # ascontiguousarray would make a 0-d array 1-d
            t = torch.from_numpy(np.array(a).view(dt)).to(device)
        else:
            t = torch.empty(shape, dtype=torch.from_numpy(np.zeros(1, dt)).dtype, device=device)
        # transport as raw bytes: every backend (nccl, gloo) moves uint8, not every one moves int16
        # (a 0-d tensor cannot be viewed as bytes: it goes through a 1-element view)
        flat = t.reshape(-1) if t.dim() == 0 else t
        _bcast(flat.view(torch.uint8) if flat.numel() else flat, src)
        out[names[i]] = t
    return out


def gather_in_order(local: np.ndarray, frame_ids: list[int], n_frames: int) -> np.ndarray | None:
    """Re-order per-rank results by sequence number on rank 0 (host side, after the D2H copies)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.empty((n_frames,) + local.shape[1:], local.dtype)
        out[frame_ids] = local
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world
    dist.all_gather_object(gathered, (frame_ids, local))
    if rank != 0:
        return None
    out = np.empty((n_frames,) + local.shape[1:], local.dtype)
    for ids, arr in gathered:
        out[ids] = arr
    return out
