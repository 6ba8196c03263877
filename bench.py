#!/usr/bin/env python
"""bench.py — CFR iterations/s of the B200 engine on BASELINE.json's config 2 (turn + river subgame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2]

One "step" = one full CFR iteration (both players traversed and updated, cfr.rs:216-226) over the whole
public tree of the workload.  N > 1 is launched by torchrun, one rank per GPU; the river boards are
sharded across ranks and the counterfactual values at the shared chance nodes are all-reduced over
NCCL every traversal (strong scaling: the job is the same subgame at every N).

Rank 0 prints ONE JSON line (keys documented in the task contract): value = iterations/s with all
inputs resident in HBM (CUDA events around each step, L2 flushed between steps, max over ranks),
e2e = the same through the C ABI from host buffers, roofline = the dominant kernel (task_kernel<CFR>,
one launch per player traversal) from CUDA events on the engine's stream, cpu_baseline = the literal
scalar port of the reference's cfr() on this box's host cores.
`--impl reference` times that port alone (the reference itself cannot be built: no Rust toolchain).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "cfr_iterations_per_sec"
UNIT = "iter/s"
BYTES_PER_UPDATE = 20  # regret R+W (8) + strategy_sum R+W (8) + opponent-side regret read (4)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(name: str):
    from rustsolver_b200 import configs
    w = getattr(configs, name)()
    import rustsolver_b200 as rb
    n_actions, tree = rb.build_game_tree(w.options)
    ranges = configs.workload_ranges(w)
    return w, tree, ranges


# ------------------------------------------------------------------------------------------------
# reference arm: the literal scalar port of cfr() on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_rate(w, tree, ranges, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times oracle.literal_cfr on a bounded sample of the root hole-card combos.

    One step = cfr(player 0) + cfr(player 1) over every `stride`-th root combo (cfr.rs:493-499 iterates
    all of them); a full iteration costs stride x that, so iterations/s = 1 / (stride * t_step)."""
    from oracle import OracleGame
    og = OracleGame(tree, ranges, w.options.board_mask if not w.board_masks else w.board_masks[0])
    n_combos = int(og.n_combos)
    # calibrate: time a thin sample, then size the stride for ~target_seconds per step
    stride = max(1, n_combos // 2048)
    t0 = time.perf_counter()
    visited, _ = og.literal_cfr(1, stride, 0)
    t_cal = time.perf_counter() - t0
    per_combo = t_cal / max(visited, 1)
    want = max(1, int(target_seconds / max(per_combo, 1e-9)))
    stride = max(1, n_combos // want)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        visited, _ = og.literal_cfr(1, stride, i % stride)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t_step = float(np.mean(times))
    iters_per_s = 1.0 / (t_step * stride)
    sample = (f"literal cfr() port, every {stride}-th of {n_combos} root hole-card combos per step "
              f"({visited} combos, {t_step:.2f} s/step), extrapolated x{stride}")
    return iters_per_s, t_step, og.num_threads, sample, og.updates_per_iter


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, tree, ranges = build_workload(args.workload)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes whatever K and W are
    target = min(8.0, 150.0 / (steps + warmup))
    rate, t_step, cores, sample, upd = cpu_port_rate(w, tree, ranges, target, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1000.0 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "i32 fixed-point x10000 (f32 math)", "data": "synthetic",
        "config": {"workload": w.name, "updates_per_iteration": int(upd)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "updates_per_sec": rate * upd,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rustsolver_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.tensor(list(rb.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    w, tree, ranges = build_workload(args.workload)
    t0 = time.perf_counter()
    eng = rb.Engine(tree, ranges, w.options.board_mask, w.card_abs, board_masks=w.board_masks, device=local_rank,
                    rank=rank, world_size=world, nccl_id=nccl_id, flags=int(os.environ.get("RS_ENGINE_FLAGS", "0")))
    fused = False
    if world > 1 and os.environ.get("RS_NO_FUSED", "0") != "1" and eng.stats().n_rounds > 1:
        # the kernel exchanges the chance-node sums itself over NVLink peer memory (False: some rank could not map
        # its peers, every rank stays on the NCCL all-reduce)
        fused = eng.enable_fused_exchange(dist, dev)
    create_s = time.perf_counter() - t0
    st = eng.stats()
    upd_global = int(st.updates_per_iteration_global)
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        flush_buf.add_(1)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K steps, each bracketed by the engine's own CUDA events ----
    for _ in range(warmup):
        eng.iterate(1)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ms_before = eng.stats().device_ms
    launches_before = eng.stats().kernel_launches
    for _ in range(steps):
        flush_l2()
        barrier()
        eng.iterate(1)  # rs_iterate records events on its launch stream around the graph replay
    barrier()
    wall = time.perf_counter() - wall0
    s1 = eng.stats()
    dev_ms = s1.device_ms - ms_before
    launches = int(s1.kernel_launches - launches_before)
    clocks = sampler.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / steps
    value = 1000.0 / ms_per_step

    # ---- back-to-back (no flush) for context ----
    barrier()
    ms_b = eng.stats().device_ms
    eng.iterate(steps)
    t = torch.tensor([eng.stats().device_ms - ms_b], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    back_to_back = 1000.0 * steps / float(t.item())

    # ---- end to end through the C ABI with host buffers ----
    # per step: H2D of both players' range weights (the host input of a re-solve step) from pinned
    # memory, rs_iterate(1) (synchronous), D2H of both players' root counterfactual values
    pinned = [torch.ones(len(ranges[p]), dtype=torch.float32).pin_memory() for p in range(2)]
    h2d = sum(x.numel() * 4 for x in pinned)
    d2h = 0
    lib = eng._lib
    e2e_t = 0.0
    for i in range(warmup + steps):
        flush_l2()
        barrier()
        t0 = time.perf_counter()
        for p in range(2):
            rc = lib.rs_set_range_weights(eng._h, p, ctypes.cast(pinned[p].data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                          pinned[p].numel())
            assert rc == 0, lib.rs_last_error()
        eng.iterate(1)
        outs = [eng.root_values(p) for p in range(2)]
        dt = time.perf_counter() - t0
        if i >= warmup:
            e2e_t += dt
        d2h = sum(o.nbytes for o in outs)
    t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = steps / float(t.item())

    # ---- roofline of the dominant kernel: CUDA events on the engine's stream, launch by launch ----
    peak, peak_src = load_peaks()
    prof_runs = []
    for _ in range(3):
        flush_l2()
        prof_runs.append(eng.profile_iteration())
    dom = [[k for k in run if k["kind"] == "traversal" and k["phase"] == 0] for run in prof_runs]
    dom_ms = float(np.mean([k["ms"] for run in dom for k in run]))
    dom_bytes = float(np.mean([k["table_bytes"] for run in dom for k in run]))
    dom_vec = float(np.mean([k["vector_bytes"] for run in dom for k in run]))
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    total_ms = float(np.mean([sum(k["ms"] for k in run) for run in prof_runs]))
    shares = {}
    for k in prof_runs[-1]:
        key = f'{k["kind"]}_p{k["traverser"]}_phase{k["phase"]}'
        shares[key] = shares.get(key, 0.0) + k["ms"]
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "task_kernel<CFR>: persistent dataflow kernel, one launch = one player's traversal of the whole tree (mean of both players)",
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                "l2_vector_bytes_per_launch": dom_vec,
                "kernel_share_of_iteration": sum(k["ms"] for k in dom[-1]) / sum(k["ms"] for k in prof_runs[-1]),
                "per_kernel_ms": {k: round(v, 5) for k, v in shares.items()},
                "whole_iteration_frac": (upd_global * BYTES_PER_UPDATE / world) / (ms_per_step * 1e-3) / 1e9 / peak}

    # exploitability of the average strategy after everything run so far (rs_best_response: two best-response
    # traversals, same kernel; parity of this number against the oracle is tests/test_gpu_parity.py's job)
    st_end = eng.stats()
    br = eng.best_response()
    exploit = {"iterations": int(st_end.iterations), "chips": 0.5 * (br[0] + br[1]),
               "best_response_values": [br[0], br[1]], "starting_pot_chips": int(w.options.starting_pot)}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, t_step, cores, sample, _ = cpu_port_rate(w, tree, ranges, target_seconds=12.0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "nodes": int(tree.n_nodes), "action_nodes": int(tree.n_actions),
                       "boards": [int(st.n_boards[k]) for k in range(st.n_rounds)],
                       "hands": [int(st.n_hands[0]), int(st.n_hands[1])],
                       "updates_per_iteration": upd_global,
                       "table_bytes_per_gpu": int(st.table_bytes),
                       "parallelism": "single GPU" if world == 1 else (
                           f"boards sharded over {world} GPUs, chance-node sums exchanged inside the traversal kernel over NVLink peer memory"
                           if fused else f"boards sharded over {world} GPUs + NCCL all-reduce at the chance nodes"),
                       "l2": "flushed between timed steps (256 MiB device write); each step timed by CUDA events on the launch stream",
                       "threads_per_block": "4 hands per thread + 1 dispatcher warp (320 for 1128 hands)"},
            "updates_per_sec": value * upd_global,
            "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "per step: rs_set_range_weights x2 from pinned host memory, rs_iterate(1), rs_root_values x2"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "exploitability": exploit,
            "clocks": clocks,
            "back_to_back_iter_per_sec": back_to_back,
            "engine_create_s": create_s,
            "wall_s_timed_region": wall,
        }
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config5"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
