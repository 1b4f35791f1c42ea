"""Host-side logic of the multi-GPU path, on CPU: the id-range partition, per-rank slicing of a swarm,
and the torch.distributed plumbing (magics_b200/dist.py) under a world_size-2 `gloo` group.
No compute calls: the engine itself needs a GPU (tests/test_gpu_shards.py, tests/test_gpu_nccl.py)."""
import os
import socket

import numpy as np
import pytest

from magics_b200 import scenarios
from magics_b200.sharded import partition


def test_partition_is_contiguous_and_balanced():
    assert partition(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert partition(3, 4).tolist() == [0, 1, 2, 3, 3]
    for n in (0, 1, 7, 1000, 1_000_000):
        for ws in (1, 2, 3, 8):
            b = partition(n, ws)
            assert b[0] == 0 and b[-1] == n and b.shape[0] == ws + 1
            d = np.diff(b)
            assert d.min() >= 0 and d.max() - d.min() <= 1


def test_lattice_rank_slices_tile_the_global_swarm():
    """bench.py builds only the rows a rank owns: the pieces must be the global lattice, bit for bit."""
    whole = scenarios.lattice(12, 9)
    for ws in (2, 3):
        b = partition(9, ws)
        parts = [scenarios.lattice(12, 9, rows=(int(b[q]), int(b[q + 1]))) for q in range(ws)]
        assert sum(p.n for p in parts) == whole.n
        for key in ("radii", "init_means", "positions"):
            assert np.array_equal(np.concatenate([getattr(p, key) for p in parts]), getattr(whole, key)), key
        assert np.array_equal(np.concatenate([p.wp_xy for p in parts]), whole.wp_xy)
        lo = 0
        for p in parts:
            s = whole.slice(lo, lo + p.n)
            assert np.array_equal(s.wp_offsets, p.wp_offsets) and np.array_equal(s.positions, p.positions)
            lo += p.n


def test_state_hash_does_not_depend_on_the_partition():
    """bench.py prints the digests at N = 1, 2, 4, 8: shard sums must equal the single-shard value, and one
    changed bit of a mean, one changed neighbour or robot_number must change them."""
    from magics_b200.sharded import state_hash

    rng = np.random.default_rng(3)
    n, V = 61, 7
    means = rng.normal(size=(n, V, 4))
    deg = rng.integers(0, 5, n)
    off = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    nb = rng.integers(0, n, off[-1]).astype(np.int32)
    rn = rng.integers(0, 10_000, off[-1]).astype(np.int64)
    whole = state_hash(0, means, off, nb, rn)
    for ws in (2, 4, 8):
        b = partition(n, ws)
        tot = [0, 0]
        for q in range(ws):
            lo, hi = int(b[q]), int(b[q + 1])
            h = state_hash(lo, means[lo:hi], off[lo:hi + 1] - off[lo], nb[off[lo]:off[hi]], rn[off[lo]:off[hi]])
            tot = [(tot[k] + h[k]) % 2 ** 64 for k in range(2)]
        assert tuple(tot) == whole, ws
    m2 = means.copy()
    m2[17, 3, 2] = np.nextafter(m2[17, 3, 2], np.inf)
    assert state_hash(0, m2, off, nb, rn)[0] != whole[0]
    m3 = means.copy()
    m3[[4, 5]] = m3[[5, 4]]  # two robots swapped: the digest is keyed by robot id
    assert state_hash(0, m3, off, nb, rn)[0] != whole[0]
    rn2 = rn.copy()
    rn2[0] += 1
    assert state_hash(0, means, off, nb, rn2)[1] != whole[1]
    assert state_hash(0, means[:0], off[:1], nb[:0], rn[:0]) == (0, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, ws, port, q):
    import torch.distributed as dist

    from magics_b200.dist import broadcast_bytes, gather_arrays, max_over_ranks, sum_over_ranks, sum_u64_over_ranks

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        payload = bytes(range(128)) if rank == 0 else None
        got = broadcast_bytes(payload, 128)
        assert got == bytes(range(128))
        assert max_over_ranks(10.0 + rank) == 10.0 + ws - 1
        assert sum_over_ranks(1.5) == 1.5 * ws
        # exact modulo-2**64 sum (bench.py's cross-N state hash): wraps, keeps every bit
        big = 0xFFFFFFFFFFFFFFF0 + rank
        assert sum_u64_over_ranks(big) == sum(0xFFFFFFFFFFFFFFF0 + r for r in range(ws)) % 2 ** 64
        # ragged per-rank arrays (a rank may own no robots)
        mine = {"mean": np.full((rank * 3, 2, 4), float(rank)), "nb": np.arange(rank * 5, dtype=np.int32)}
        parts = gather_arrays(mine, rank, ws)
        if rank == 0:
            assert len(parts) == ws
            for r, p in enumerate(parts):
                assert p["mean"].shape == (r * 3, 2, 4) and np.all(p["mean"] == r)
                assert p["nb"].tolist() == list(range(r * 5)) and p["nb"].dtype == np.int32
        else:
            assert parts is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - reported to the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dist_helpers_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == {0: "ok", 1: "ok"}, res
