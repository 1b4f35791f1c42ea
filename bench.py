#!/usr/bin/env python
"""bench.py — robot-GBP-iterations/s of the GBP iterate hot path on B200.

A "step" is one simulation tick of the hot path over the whole synthetic swarm:
neighbour search / InterRobot factor maintenance, the two prior updates, and
`iterate_gbp_v2` with the config's schedule (10 internal + 10 external,
interleave-evenly => 10 sub-steps, each one robot-GBP-iteration per robot).

  value  robots x sub-steps x steps / device time, swarm resident in HBM
  e2e    same through the C ABI with HOST buffers each step (comms mask and
         waypoint indices in, every variable's mean out)
  roofline  dominant kernel k_iterate_axis<EXT,INT>: compulsory bytes of the
            store layout per launch / its average CUDA-event
            duration, against MEASURED_PEAKS.json; `traffic` = ncu DRAM bytes
  cpu_baseline  the oracle (C++ restatement of the reference algorithm) timed
            on the host cores on a bounded sample of the same workload

`--impl reference` times that CPU restatement alone (the Rust reference cannot be
built in this image: no rustc/cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "robot_gbp_iterations_per_sec"
UNIT = "robot-GBP-iterations/s"


def reference_message_bytes_per_robot_iteration(V: int, K: float, e_obs: int, e_trk: int) -> float:
    """what the REFERENCE's formulation moves — every message a 192-byte payload in an inbox.
    Reported for context only; the engine never materialises these messages."""
    E = 2 * (V - 1) + e_obs * (V - 2) + e_trk * (V - 2)
    F = K * (V - 1)
    return 192.0 * (5 * E + 7 * F) + 704.0 * V


def compulsory_bytes_per_robot_iteration(V: int, K: float) -> float:
    """Bytes one fused external+internal half-step of k_iterate_axis has to move for one robot with the store of
    DESIGN.md section 3 (decoupled regime, both lanes of every variable; derivation in DESIGN.md section 4):
      read   per variable: record rows of both axes 12 + mean 4, stored Dynamic messages 2 x 12, prior 5 doubles,
             record epoch 4 B; variables >= 1: last delivered mean 2 doubles
      write  per variable: new record 16 doubles + epoch 4 B + lazy-covariance flag 1 B; new Dynamic messages
             12 doubles for each of the 2 (V - 1) factor slots; variables >= 1: delivered mean 2 doubles
      per robot: mode / idle / antenna / latest bytes, iteration count, edge range, nlow (25 B)
      per edge:  neighbour slot 4, birth epoch 4, frozen bit 1, safety distance 8 B; the neighbours' position
             means are rows of records the launch reads anyway (L2 serves the K re-reads).
    """
    reads = V * (16 + 24 + 5) * 8 + V * 4 + (V - 1) * 2 * 8
    writes = V * 16 * 8 + V * 5 + 2 * (V - 1) * 12 * 8 + (V - 1) * 2 * 8
    return float(reads + writes + 25 + 17.0 * K)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def _nvml_loop(self):
        import pynvml as nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.halt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.nvml_rows.append((float(sm), [k for k, b in names.items() if bits & b]))
            except Exception:
                pass
            self.halt.wait(0.004)

    def start(self):
        # NVML in-process (a sample every 4 ms: multi-GPU timed regions last tens of milliseconds); nvidia-smi -lms as
        # the fallback when the binding is missing
        self.nvml_rows, self.handle = [], None
        try:
            import pynvml as nv
            nv.nvmlInit()
            phys = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    phys = self.index
            self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.halt = threading.Event()
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            self.proc = None
            return
        except Exception:
            self.handle = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if getattr(self, "handle", None) is not None:
            self.halt.set()
            self.t.join(timeout=1)
            sm = [r[0] for r in self.nvml_rows]
            reasons = sorted({x for r in self.nvml_rows for x in r[1]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                    "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(workload_name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_iterate<EXT,INT> launch from the committed
    `ncu --set full` capture of this workload (profiles/iterate_traffic.json), or (None, None)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "iterate_traffic.json")))[workload_name]
        return float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        return None, None


def build_workload(name: str, robots: int, rank: int = 0, world: int = 1):
    """The robots rank `rank` of `world` owns (a contiguous id range of the global swarm), and the
    global robot count.  lattice: horizontal slabs of rows (BASELINE config 5); rings: arcs of
    consecutive robots (config 4)."""
    from magics_b200 import scenarios
    from magics_b200.sharded import partition

    if name == "rings":
        sw = scenarios.rings(robots)
        b = partition(sw.n, world)
        return (sw if world == 1 else sw.slice(int(b[rank]), int(b[rank + 1]))), sw.n
    if name == "lattice":
        side = int(round(robots ** 0.5))
        ny = max(1, robots // side)
        b = partition(ny, world)
        return scenarios.lattice(side, ny, rows=(int(b[rank]), int(b[rank + 1]))), side * ny
    if name == "dense":
        # stress workload (scenarios.dense_lattice): active InterRobot factors, edges churning; slabs of rows like lattice
        side = int(round(robots ** 0.5))
        sw = scenarios.dense_lattice(side, max(1, robots // side))
        b = partition(sw.n // side, world)
        return (sw if world == 1 else sw.slice(int(b[rank]) * side, int(b[rank + 1]) * side)), sw.n
    raise ValueError(name)


def run_cpu(sw, threads: int, ticks: int):
    """Oracle timing: whole ticks, wall clock (best of `ticks` after one warm-up tick)."""
    from oracle.oracle import OracleWorld, schedule

    o = OracleWorld(sw.cfg, threads=threads)
    sw.add_to(o)
    oi, oe = schedule(sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external)
    substeps = int(np.sum(oi & oe)) or int(max(oi.sum(), oe.sum()))
    o.step()  # warm-up: creates the InterRobot factors
    best = float("inf")
    for _ in range(ticks):
        t = time.perf_counter()
        o.step()
        best = min(best, time.perf_counter() - t)
    o.close()
    return sw.n * substeps / best, best, substeps


def extra_point(sw, steps: int, warmup: int = 3) -> dict:
    """One more workload on one GPU, device-timed like the headline: robots x sub-steps x steps / CUDA-event time."""
    from magics_b200 import World, gbp_schedule

    g = World(sw.cfg, device=0)
    sw.add_to(g)
    oi, oe = gbp_schedule(sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external)
    substeps = int(np.sum(oi & oe)) or int(max(oi.sum(), oe.sum()))
    for _ in range(warmup):
        g.step()
    g.sync()
    g.set_profiling(True)
    e0 = int(g.read_connections()[0][-1])
    g.timer_start()
    for _ in range(steps):
        g.step()
    ms = g.timer_stop_ms()
    prof = g.read_profile()
    off, nbr, rn = g.read_connections()
    ax, gen = g.read_iterate_path()
    out = {"workload": sw.name, "robots": sw.n, "variables": int(sw.cfg.num_variables), "steps": steps,
           "value": sw.n * substeps * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "mean_neighbours": float(off[-1]) / max(1, sw.n), "edges_before": e0, "edges_after": int(off[-1]),
           "robots_in_k_iterate_axis": ax, "robots_in_k_iterate": gen,
           "profile_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in prof.items() if v["count"]}}
    g.close()
    return out


def scenario_point(name: str, ticks: int, cores: int) -> dict:
    """A BASELINE scenario (configs 1-3) at its real size on the reference's own inputs (tests/golden/scenarios.json):
    robots spawned by the scenario's formation clock, ms per sim step of the engine (wall clock around every
    gbp_world_step incl. a device sync: these swarms are launch-latency bound) and of the CPU restatement."""
    from magics_b200 import World, scenarios
    from oracle import oracle as _oracle
    from oracle.oracle import OracleWorld

    sc = scenarios.ReferenceScenario(name)
    # g: the default path (a launch per half-step); f: the same world with the whole iterate_gbp in one cooperative
    # launch (gbp_world_set_iterate_path(w, 2)); o: the CPU restatement
    g, f, o = World(sc.cfg, device=0), World(sc.cfg, device=0), OracleWorld(sc.cfg, threads=cores)
    f.set_single_launch_tick(True)
    g.set_sdf_from_environment(sc.env)
    f.set_sdf_from_environment(sc.env)
    o.set_sdf(_oracle.env_to_sdf_image(sc.env))
    rng = np.random.default_rng(0)
    events = sc.spawn_events(ticks)
    t_gpu = t_fused = t_cpu = 0.0
    timed = 0
    for tick in range(ticks):
        for _, k in [e for e in events if e[0] == tick]:
            sw = sc.spawn(k, rng)
            if sw is not None:
                for w in (g, f, o):
                    sw.add_to(w, set_sdf=False)
        if g.num_robots == 0:
            continue
        for w in (g, f):
            w.sync()
            t = time.perf_counter()
            w.step()
            w.sync()
            if w is g:
                t_gpu += time.perf_counter() - t
            else:
                t_fused += time.perf_counter() - t
        t = time.perf_counter()
        o.step()
        t_cpu += time.perf_counter() - t
        timed += 1
    n = g.num_robots
    edges = int(g.read_connections()[0][-1])
    same = bool(np.array_equal(g.read_beliefs()["mean"], f.read_beliefs()["mean"], equal_nan=True))
    g.close()
    f.close()
    o.close()
    return {"scenario": name, "robots_at_end": n, "variables": int(sc.cfg.num_variables), "ticks_timed": timed,
            "edges_at_end": edges, "ms_per_step_gpu": t_gpu * 1e3 / max(1, timed),
            "ms_per_step_gpu_single_launch": t_fused * 1e3 / max(1, timed), "single_launch_same_bits": same,
            "ms_per_step_cpu_port": t_cpu * 1e3 / max(1, timed), "cpu_threads": cores}


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version
    # banner under NCCL_DEBUG, torchrun notices) is sent to stderr while the benchmark runs.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="lattice", choices=["rings", "lattice", "dense"])
    ap.add_argument("--robots", type=int, default=None,
                    help="robots of the whole swarm (default: lattice 1 000 000, rings 100 000)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --robots is the whole swarm at every N (BASELINE config 5: 1M robots over "
                         "1/2/4/8 GPUs); weak: --robots per GPU")
    ap.add_argument("--cpu-robots", type=int, default=3000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra workloads reported next to the headline at N = 1 (config 4 rings-100k, a 10 k "
                         "lattice, the dense stress lattice, BASELINE scenarios 1-3 at their real sizes)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.robots is None:
        args.robots = {"lattice": 1_000_000, "rings": 100_000, "dense": 250_000}[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        # The reference's own CPU implementation of the path = the oracle's C++ restatement of the
        # Rust algorithm (no Rust toolchain here), all host threads, on a bounded
        # sample of the same workload: W warm-up ticks, then exactly K timed ticks.
        if rank != 0:
            return
        from magics_b200 import scenarios
        from oracle import oracle as _oracle
        from oracle.oracle import OracleWorld, schedule

        # this arm must not map the product library: the generators take their variable timesteps from the
        # oracle's restatement of utils.rs:95-133 instead of the engine's gbp_variable_timesteps
        scenarios.get_variable_timesteps = _oracle.variable_timesteps
        sw, _ = build_workload(args.workload, args.cpu_robots)
        assert "libgbp_b200" not in open("/proc/self/maps").read(), "reference arm loaded the product library"
        o = OracleWorld(sw.cfg, threads=cores)
        sw.add_to(o)
        oi, oe = schedule(sw.cfg.schedule_kind, sw.cfg.iterations_internal, sw.cfg.iterations_external)
        substeps = int(np.sum(oi & oe))
        for _ in range(max(1, args.warmup)):
            o.step()
        t = time.perf_counter()
        for _ in range(args.steps):
            o.step()
        secs = time.perf_counter() - t
        o.close()
        v = sw.n * substeps * args.steps / secs
        sample = (f"{sw.n} robots of the {args.workload} workload ({sw.name}), {args.steps} sim ticks of {substeps} "
                  f"sub-steps after {max(1, args.warmup)} warm-up; C++ restatement of the reference algorithm "
                  "(Rust toolchain absent), threads over robots for the internal half, serial external half")
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}-{args.robots}: V={int(sw.cfg.num_variables)}, "
                                   "dyn+obstacle+interrobot factors, interleave-evenly 10/10; CPU arm runs the "
                                   f"bounded sample {sw.name}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return

    import torch
    import torch.distributed as dist

    from magics_b200 import World

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from magics_b200 import gbp_schedule, pinned_empty
    from magics_b200.dist import broadcast_comm_id, max_over_ranks, sum_over_ranks, sum_u64_over_ranks
    from magics_b200.sharded import state_hash

    # One swarm, spatially partitioned: rank q owns a contiguous id range (a slab of lattice rows /
    # an arc of rings) and exchanges the published records of its border robots with its
    # neighbours by NCCL send/recv before every external half-step.
    total_robots = args.robots * (world if args.scaling == "weak" else 1)
    sw, n_total = build_workload(args.workload, total_robots, rank, world)
    cfg = sw.cfg
    if world > 1:
        g = World.create_shard(cfg, local_rank, rank, world, broadcast_comm_id(rank))
    else:
        g = World(cfg, device=local_rank)
    sw.add_to(g)
    g.commit_shards()
    assert g.num_robots_global == n_total

    oi, oe = gbp_schedule(cfg.schedule_kind, cfg.iterations_internal, cfg.iterations_external)
    substeps = int(np.sum(oi & oe))
    n = sw.n

    def barrier():
        g.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------
    for _ in range(args.warmup):
        g.step()
    launches0 = g.kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    g.timer_start()
    for _ in range(args.steps):
        g.step()
    ms = g.timer_stop_ms()
    barrier()
    clocks = sampler.stop()
    launches = g.kernel_launches - launches0
    launches = sum_over_ranks(launches) if world > 1 else launches
    # per-kernel device times for the roofline leg: a separate, untimed pass — the CUDA events around every launch
    # would sit between kernels that otherwise overlap their launch with the previous kernel's tail
    prof_steps = min(args.steps, 5)
    g.set_profiling(True)
    for _ in range(prof_steps):
        g.step()
    prof = g.read_profile()
    g.set_profiling(False)
    ms = max_over_ranks(ms) if world > 1 else ms
    value = n_total * substeps * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers ---------------------
    # every step: comms mask + waypoint indices host -> device, one sim tick, every variable's
    # mean device -> host (what the Bevy systems / visualisers read back each tick)
    ant = pinned_empty((n,), np.uint8)
    wpi = pinned_empty((n,), np.int32)
    # two host buffers: the means of tick t travel to the host while tick t+1 runs on the device
    means = [pinned_empty((n, cfg.num_variables, 4), np.float64) for _ in range(2)]
    ant[:] = 1
    wpi[:] = 1
    for k in range(2):  # warm the read-back path (scratch allocation)
        g.read_means_into(means[k])
    checksum = 0.0
    barrier()
    t0 = time.perf_counter()
    g.timer_start()
    for k in range(args.steps):
        g.set_comms(ant, None)
        g.set_waypoint_index(wpi)
        g.step()
        g.read_means_into_async(means[k % 2])  # waits for the previous tick's copy, then starts this one
        if k > 0:
            checksum += float(means[(k - 1) % 2][-1, -1, 0])  # the host consumes the previous tick's result
    g.readback_wait()
    checksum += float(means[(args.steps - 1) % 2][-1, -1, 0]) if n else 0.0
    ms_e2e = g.timer_stop_ms()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(ms_e2e, wall_e2e)
    e2e_ms = max_over_ranks(e2e_ms) if world > 1 else e2e_ms
    e2e_value = n_total * substeps * args.steps / (e2e_ms * 1e-3)
    # after the timed ticks: digests of every variable mean and of the whole InterRobot edge set with its
    # robot_numbers — identical for N = 1, 2, 4, 8 by design (the sharded run reproduces the single-GPU bits)
    conn = g.read_connections()
    g.read_means_into(means[0])
    h_means, h_conn = state_hash(g.first_global_id, means[0], *conn)
    if world > 1:
        h_means, h_conn = sum_u64_over_ranks(h_means), sum_u64_over_ranks(h_conn)
    rn_max_local = float(conn[2].max()) if conn[2].size else -1.0
    rn_max = max_over_ranks(rn_max_local) if world > 1 else rn_max_local
    edges_local = float(conn[0][-1])
    edges_total = sum_over_ranks(edges_local) if world > 1 else edges_local
    ghosts_total = sum_over_ranks(g.num_ghosts) if world > 1 else 0.0
    h2d_total = sum_over_ranks(ant.nbytes + wpi.nbytes) if world > 1 else float(ant.nbytes + wpi.nbytes)
    d2h_total = sum_over_ranks(means[0].nbytes) if world > 1 else float(means[0].nbytes)

    if rank == 0:
        deg = float(np.diff(g.read_connections()[0]).mean()) if n else 0.0
        peak, peak_kind = measured_peak_gbs()
        bytes_iter = compulsory_bytes_per_robot_iteration(cfg.num_variables, deg)
        dom = prof.get("iterate_ext_int", {"count": 0, "ms": 0.0})
        roof = None
        if dom["count"]:
            # device time of one fused external+internal half-step pair; a sharded run launches the kernel twice per
            # pair (border robots, then the rest), so the total is divided by the pairs of the schedule
            ph = [c for a, b in zip(oi, oe) for c in (("I",) if a else ()) + (("E",) if b else ())]
            fused_per_tick = sum(1 for k in range(len(ph) - 1) if ph[k] == "E" and ph[k + 1] == "I")
            avg_s = dom["ms"] * 1e-3 / max(1, fused_per_tick * prof_steps)
            achieved = bytes_iter * n / avg_s / 1e9
            traffic, traffic_src = ncu_traffic(sw.name)
            if traffic is not None and n_total != n:
                traffic *= n / n_total  # the ncu capture is of the whole swarm on one GPU; this shard iterates n robots
            roof = {"bound": "hbm", "kernel": "k_iterate_axis<EXT,INT>", "achieved": achieved, "peak": peak,
                    "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "traffic_frac": (traffic / avg_s / 1e9 / peak) if traffic else None,
                    "avg_launch_ms": avg_s * 1e3, "launches_timed": dom["count"],
                    "compulsory_bytes_per_robot_iteration": bytes_iter, "mean_neighbours": deg,
                    "reference_message_bytes_per_robot_iteration": reference_message_bytes_per_robot_iteration(
                        cfg.num_variables, deg, int(cfg.enable_obstacle), int(cfg.enable_tracking)),
                    "note": "achieved = compulsory bytes of this store layout (DESIGN.md section 4) x robots of "
                            "rank 0 / average device time of the fused external+internal launch (CUDA events "
                            "on the engine's stream); traffic = dram__bytes_read + dram__bytes_write of the same "
                            "launch from the committed ncu capture (scaled to this rank's robots when sharded); "
                            "traffic_frac = traffic / time / peak"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is an N = 1 figure (rank 0 alone would hold up the others)
            swc, _ = build_workload(args.workload, args.cpu_robots)
            v, secs, _ = run_cpu(swc, cores, 2)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{swc.n} robots of the {args.workload} workload, 1 sim tick ({substeps} sub-steps), "
                             f"best of 2 after warm-up, {secs * 1e3:.1f} ms; C++ restatement of the reference "
                             "algorithm (Rust toolchain absent)"}
            try:  # "also report the 1-thread figure" - a quarter of the sample, one timed tick
                sw1, _ = build_workload(args.workload, max(64, args.cpu_robots // 4))
                v1, secs1, _ = run_cpu(sw1, 1, 1)
                cpu["value_1_thread"] = v1
                cpu["sample_1_thread"] = f"{sw1.n} robots, 1 sim tick after warm-up, {secs1 * 1e3:.1f} ms, 1 thread"
            except Exception as e:  # the extra figure must never cost the bench line
                cpu["value_1_thread"] = None
                cpu["sample_1_thread"] = f"failed: {e}"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{sw.name}: {n_total} robots x V={cfg.num_variables}, mean K="
                                   f"{edges_total / max(1, n_total):.2f}, dyn+obstacle+interrobot factors, "
                                   "interleave-evenly 10/10",
                       "robots_total": n_total, "robots_rank0": n, "sub_steps_per_step": substeps,
                       "partition": (f"{world} shards of contiguous robot ids (lattice: slabs of rows), "
                                     f"{int(ghosts_total)} ghost robots in total, NCCL send/recv halo before every "
                                     "external half-step") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (store >> 126 MB)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h_total),
                    "what": "per step: set_comms + set_waypoint_index (pinned host -> device), gbp_world_step, "
                            "all variable means device -> pinned host (copy of tick t overlaps tick t+1, consumed one tick "
                            "later; the last copy is inside the timed region); max(CUDA events, wall clock), max over ranks"},
            "state_hash": {"means": f"{h_means:016x}", "connectivity": f"{h_conn:016x}", "edges": int(edges_total),
                           "last_robot_number": int(rn_max), "ticks": args.warmup + 2 * args.steps + prof_steps,
                           "what": "sum mod 2^64 over all robots of a 64-bit digest of every variable mean "
                                   "(keyed by global robot id) / of every directed InterRobot pair with its "
                                   "robot_number, after all ticks of this run; the same for every N"},
            "roofline": roof, "cpu_baseline": cpu, "profile_ms": prof, "profile_steps": prof_steps,
        }
        if world == 1 and not args.no_extras and args.workload == "lattice":
            # the other sizes the metric names (10 k / 100 k), the regime the headline never enters (active
            # InterRobot factors, edges created and deleted every tick) and configs 1-3 on the reference's inputs
            from magics_b200 import scenarios

            g.close()
            extras = []
            for make, steps in ((lambda: scenarios.rings(100_000), 10), (lambda: scenarios.lattice(100, 100), 20),
                                (lambda: scenarios.dense_lattice(500, 500), 5)):
                try:
                    extras.append(extra_point(make(), steps))
                except Exception as e:  # an extra never takes the headline down
                    extras.append({"error": repr(e)})
            line["points"] = extras
            scen = []
            for name, ticks in (("Circle Experiment", 30), ("Structured Junction Twoway", 70),
                                ("Collaborative Complex", 50)):
                try:
                    scen.append(scenario_point(name, ticks, cores))
                except Exception as e:
                    scen.append({"scenario": name, "error": repr(e)})
            line["scenarios"] = scen
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
