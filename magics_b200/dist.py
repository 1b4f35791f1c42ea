"""torch.distributed plumbing for one-process-per-GPU runs (bench.py, tests): nothing here is on the
compute path — the engine's own halo exchange is NCCL send/recv inside the C library.  Every helper
works on any initialised backend (`nccl` on the GPU box, `gloo` in the CPU tests)."""
from __future__ import annotations

import io

import numpy as np


def _device():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Fixed-size byte string from rank `src` to every rank."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(nbytes, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        if payload is None or len(payload) != nbytes:
            raise ValueError("broadcast_bytes: the source rank must pass exactly nbytes bytes")
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def broadcast_comm_id(rank: int) -> bytes:
    """Rank 0 creates the engine's NCCL communicator id (gbp_comm_unique_id) and every rank receives it."""
    from .world import COMM_ID_BYTES, comm_unique_id

    return broadcast_bytes(comm_unique_id() if rank == 0 else None, COMM_ID_BYTES, src=0)


def max_over_ranks(value: float) -> float:
    """Device time of a multi-GPU region = the slowest rank's."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def sum_u64_over_ranks(value: int) -> int:
    """Sum modulo 2**64 of one unsigned 64-bit integer per rank, exact on every backend (four 16-bit limbs
    summed as int64)."""
    import torch
    import torch.distributed as dist

    v = int(value) & 0xFFFFFFFFFFFFFFFF
    t = torch.tensor([(v >> (16 * k)) & 0xFFFF for k in range(4)], dtype=torch.int64, device=_device())
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    limbs = [int(x) for x in t.cpu().tolist()]
    return sum(l << (16 * k) for k, l in enumerate(limbs)) & 0xFFFFFFFFFFFFFFFF


def gather_arrays(arrays: dict, rank: int, world_size: int, dst: int = 0):
    """Dict of numpy arrays from every rank to `dst` (list indexed by rank; None elsewhere).
    Arrays may differ in length per rank: each rank's dict travels as one npz byte string."""
    import torch
    import torch.distributed as dist

    buf = io.BytesIO()
    np.savez(buf, **{k: np.asarray(v) for k, v in arrays.items()})
    raw = np.frombuffer(buf.getvalue(), np.uint8)
    dev = _device()
    sizes = torch.zeros(world_size, dtype=torch.int64, device=dev)
    sizes[rank] = raw.size
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    cap = int(sizes.max().item())
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    mine[: raw.size] = torch.from_numpy(raw.copy()).to(dev)
    out = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world_size)]
    dist.all_gather(out, mine)
    if rank != dst:
        return None
    res = []
    for q in range(world_size):
        b = out[q][: int(sizes[q].item())].cpu().numpy().tobytes()
        z = np.load(io.BytesIO(b))
        res.append({k: z[k] for k in z.files})
    return res
