"""Multi-process check of the NCCL transport (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/nccl_shard_check.py [scenario]

Every rank owns a contiguous slice of one swarm (`World.create_shard`, NCCL send/recv halo), steps it,
and rank 0 gathers all shards' beliefs / connectivity and compares them
  * bit for bit with the single-GPU engine run on rank 0's GPU, and
  * for the small scenario, within 1e-9 with the CPU oracle.
Prints `NCCL-SHARDS-OK ...` on success; any mismatch raises."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from magics_b200 import World, scenarios
    from magics_b200.dist import broadcast_comm_id, gather_arrays
    from magics_b200.sharded import partition

    rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = sys.argv[1] if len(sys.argv) > 1 else "circle"
    late = None
    if which == "late":
        # robots spawned while the simulation runs: the last rank takes them, every rank commits again
        sw, ticks = scenarios.circle(16, 14.0), 24
        late = scenarios.circle(9, 11.0)
    elif which == "circle":
        sw, ticks = scenarios.circle(30), 25
    elif which == "lattice":
        sw, ticks = scenarios.lattice(60, 40), 4
    else:
        sw, ticks = scenarios.rings(20000), 3
    b = partition(sw.n, ws)
    mine = sw.slice(int(b[rank]), int(b[rank + 1]))
    g = World.create_shard(sw.cfg, local, rank, ws, broadcast_comm_id(rank))
    mine.add_to(g)
    g.commit_shards()
    assert g.first_global_id == int(b[rank]) and g.num_robots_global == sw.n
    for tick in range(ticks):
        g.step()
        if late is not None and tick == 8:
            if rank == ws - 1:
                late.add_to(g, set_sdf=False)
            g.commit_shards()
            assert g.num_robots_global == sw.n + late.n
    bel = g.read_beliefs()
    off, nb, rn = g.read_connections()
    parts = gather_arrays({**bel, "deg": np.diff(off), "nb": nb, "rn": rn, "ghosts": np.array([g.num_ghosts])}, rank, ws)
    if rank == 0:
        got = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
        one = World(sw.cfg, device=local)
        sw.add_to(one)
        for tick in range(ticks):
            one.step()
            if late is not None and tick == 8:
                late.add_to(one, set_sdf=False)
        ref = one.read_beliefs()
        for k in ("eta", "lam", "mean", "cov", "valid"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), f"{which}: {k} differs from the single-GPU engine"
        o1, n1, r1 = one.read_connections()
        assert np.array_equal(got["deg"], np.diff(o1)) and np.array_equal(got["nb"], n1) and np.array_equal(got["rn"], r1)
        assert got["ghosts"].sum() > 0
        if which in ("circle", "late"):
            from oracle.oracle import OracleWorld
            from tests.parity import assert_beliefs_match

            o = OracleWorld(sw.cfg)
            sw.add_to(o)
            for tick in range(ticks):
                o.step()
                if late is not None and tick == 8:
                    late.add_to(o, set_sdf=False)
            assert_beliefs_match(got, o.read_beliefs(), what="nccl shards vs oracle")
            oo, no, ro = o.read_connections()
            assert np.array_equal(got["nb"], no) and np.array_equal(got["rn"], ro)
        print(f"NCCL-SHARDS-OK {which} ws={ws} robots={sw.n + (late.n if late else 0)} ghosts={got['ghosts'].tolist()}", flush=True)
    dist.barrier()
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
