#!/usr/bin/env python
"""bench.py -- 9x9x6 env-steps/sec (cascades included) of the batched stepping engine.

    python bench.py --gpus N --steps K --warmup W [--scaling weak|strong] [--impl reference]

Workload = BASELINE.json configs[2] (SURVEY.md 8d "config 3"): lockstep 9x9x6 boards, Philox keyed by global board
index (results do not depend on N), each step = pick a uniformly random legal action from the previous legal mask ->
swap -> full cascade -> reward/done/won -> new legal mask.  One step of all boards is ONE ecg_step call (two kernel
launches).  Inputs are resident in HBM (1.1 GB of boards + masks per 2^24 boards, far larger than the 126 MB L2, so
no explicit L2 flush is needed between steps).

--scaling weak (default): 2^24 boards PER GPU, `value` is that run; the JSON line also carries "strong" = the same
K steps on 2^24 boards IN TOTAL split evenly over the N GPUs (BASELINE configs[2] "16M boards ... sharded over
1/2/4/8").  --scaling strong swaps the two.

The JSON line also carries
  sustained    >= 5 s of back-to-back steps with SM clock and power sampled over them (the K-step region is 0.05 s);
  shape_sweep  BASELINE configs[3]: 6x6x4, 9x9x6, 12x12x7, 16x16x8 at 2^22 boards per GPU, roofline fraction per shape;
  mcts         BASELINE configs[4]: standard MCTS with 2^20 GPU rollouts per simulation, reward sums over NCCL;
  replay       the REFERENCE's dynamics at the same size: ECG_REFILL_REPLAY, every step restarts the MT19937 stream
               of cfg.seed (boardv2.py:46) -- one stream shared by all boards, and 4096 distinct streams;
  e2e          the same metric through the public host-buffer API (HostStepper): per step the actions are fetched to
               pinned host memory (board.random_action()), copied back H2D, and obs/reward/done/won are read D2H, all
               inside the timed region; "e2e" is the env-contract form (uint8 cells in the caller's pinned buffer) by
               the faster of two transports, both reported: "e2e_uint8_direct" (the uint8 cells cross PCIe) and
               "e2e_uint8_host_expand" (4-bit codes cross PCIe, libecg.so's host threads widen them, byte-identical
               result); "e2e_nibbles" hands the caller the 4-bit form itself;
  roofline     algorithmic bytes per launch (117 B per env-step) / average step-kernel duration;
  cpu_baseline the CPU oracle port (oracle/, test infrastructure) timed on this box's host cores, rank 0, N=1.
--impl reference times that CPU port on all host threads as the reference arm (the reference itself is pure
Python with no compiled sources and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "9x9x6 env-steps/sec (cascades incl.)"
UNIT = "env-steps/s"
KEY = 0x9E3779B97F4A7C15
SHAPE = (9, 9, 6)
BYTES_PER_STEP = 117  # 2*ceil(81/2) + 17 + ceil(144/8), SURVEY.md 8d
FALLBACK_HBM_GBS = 6650.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boards", type=int, default=1 << 24, help="boards per GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-chunks", type=int, default=32, help="HostStepper chunks (copy / kernel pipeline depth)")
    ap.add_argument("--sustained-seconds", type=float, default=5.0)
    ap.add_argument("--replay-steps", type=int, default=12)
    ap.add_argument("--sweep-boards", type=int, default=1 << 22, help="boards per GPU and shape of the shape sweep")
    ap.add_argument("--mcts-leaves", type=int, default=1 << 20)
    ap.add_argument("--mcts-sims", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the strong/weak twin, sustained and replay legs")
    return ap.parse_args()


def workload(boards_per_gpu, n_gpus):
    boards_per_gpu = int(boards_per_gpu)
    return {
        "workload": "configs[2]: lockstep 9x9x6 boards, random legal action from the previous mask + step + "
                    "cascade + reward/done/won + legal-swap mask every step, Philox4x32-10 refill",
        "boards_per_gpu": boards_per_gpu, "boards_total": boards_per_gpu * n_gpus,
        "rows": 9, "cols": 9, "types": 6, "refill": "philox4x32-10", "legal_mask_each_step": True,
        "l2": "inputs (boards+masks, %.2f GB per GPU) exceed the 126 MB L2; no flush" %
              (boards_per_gpu * (48 + 20) / 1e9),
        "parallelism": f"{n_gpus} x independent shards by global board index, no collective in the step path",
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the step kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_min": min(sm), "power_w_median": statistics.median(sorted(pw)[len(pw) // 2:]) if pw else None}


# ------------------------------------------------------------------ CPU port (oracle) legs

def cpu_port_rate(target_seconds, threads=None):
    """env-steps/s of the oracle port on the host: lockstep Philox steps over a bounded sample of boards."""
    import numpy as np
    from oracle.oracle import Oracle
    o = Oracle(*SHAPE)
    if threads:
        Oracle.set_threads(threads)
    cores = Oracle.max_threads()
    n0 = 4096 * max(1, cores // 4)
    boards = np.stack([o.init_board(o.rng_philox(KEY, i, 0xFFFFFFFF)) for i in range(512)])
    boards = np.tile(boards, (n0 // 512 + 1, 1, 1))[:n0].copy()
    t = time.perf_counter()
    o.philox_episode_batch(boards, KEY, 0, 4, step0=0, inplace=True)
    rate = n0 * 4 / (time.perf_counter() - t)
    moves = 20
    n = int(min(max(rate * target_seconds / moves, 1024), 1 << 22)) // 512 * 512
    boards = np.tile(boards[:512], (n // 512, 1, 1)).copy()
    t = time.perf_counter()
    _, _, steps = o.philox_episode_batch(boards, KEY, 0, moves, step0=4, inplace=True)
    dt = time.perf_counter() - t
    return {"value": float(steps.sum() / dt), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} boards x {moves} lockstep steps (pick + step + cascade + legal mask), "
                      f"{dt:.1f} s on {cores} host threads; C oracle port of the reference's Python/NumPy path"}


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port; the reference is Python-only) on all
    host threads; each step = one lockstep step of a bounded sample of the workload's boards."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle.oracle import Oracle
    o = Oracle(*SHAPE)
    cores = Oracle.max_threads()
    n = 2048 * cores
    base = np.stack([o.init_board(o.rng_philox(KEY, i, 0xFFFFFFFF)) for i in range(512)])
    boards = np.tile(base, (n // 512 + 1, 1, 1))[:n].copy()
    t = time.perf_counter()
    o.philox_episode_batch(boards, KEY, 0, 2, step0=0, inplace=True)
    rate = n * 2 / (time.perf_counter() - t)
    budget = 120.0 / max(args.steps + args.warmup, 1)  # whole run within a few minutes
    per_step = min(0.5, budget)
    n = int(min(max(rate * per_step, 512), args.boards)) // 512 * 512
    boards = np.tile(base, (n // 512, 1, 1)).copy()
    for w in range(args.warmup):
        o.philox_episode_batch(boards, KEY, 0, 1, step0=w, inplace=True)
    t = time.perf_counter()
    done = 0
    for k in range(args.steps):
        _, _, st = o.philox_episode_batch(boards, KEY, 0, 1, step0=args.warmup + k, inplace=True)
        done += int(st.sum())
    dt = time.perf_counter() - t
    value = done / dt
    cfg = workload(args.boards, args.gpus)
    sample = f"{n} of the {args.boards} boards per step, {cores} host threads, C oracle port (reference is pure Python)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    E = importlib.import_module("element-crush-gym_b200")
    rank, world, local = E.dist.init_from_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    import torch.distributed as tdist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = E._native.lib()
    strong_main = args.scaling == "strong"
    first, count = E.dist.shard_range(args.boards, world, rank)  # the strong split of --boards
    n_main, board0_main = (count, first) if strong_main else (args.boards, rank * args.boards)

    def barrier():
        if world > 1:
            tdist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    def make_env(n, board0):
        # episodes long enough that no board turns terminal inside the run (num_moves is a free env parameter)
        env = E.BatchedMatch3Env(n, 9, 9, 6, num_moves=1 << 30, env_goal=500, seed=KEY, device=dev, refill="philox",
                                 board0=board0)
        env.board.packed_mask()
        return env

    def timed_steps(b, steps, warmup, mark_mid=False):
        """K lockstep steps bracketed by barrier + synchronize; -> (ms max over ranks, per-step ms, per-launch ms)"""
        for _ in range(warmup):
            b.apply_action(None)
        torch.cuda.synchronize(dev)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        two = mark_mid and b.two_kernel_step
        # the step is TWO kernels (common-case kernel over all boards, exact kernel over the boards it hands off);
        # ecg_step records mid[k] between them so the dominant kernel is timed by itself, live, on its own stream
        mid = [torch.cuda.Event(enable_timing=True) for _ in range(steps)] if two else []
        for e in mid:
            e.record()  # creates the cudaEvent_t handles
        torch.cuda.synchronize(dev)
        ev[0].record()
        for k in range(steps):
            if two:
                L.ecg_step_mark_event(ctypes.c_void_p(mid[k].cuda_event))
            b.apply_action(None)  # pick + swap + cascade + reward/flags + new mask
            ev[k + 1].record()
        torch.cuda.synchronize(dev)
        barrier()
        ms = ev[0].elapsed_time(ev[-1])
        per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        per_launch = [ev[k].elapsed_time(mid[k]) for k in range(steps)] if two else per_step
        return max_over_ranks(ms), per_step, per_launch

    # ---- the main leg
    env = make_env(n_main, board0_main)
    b = env.board
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.ecg_launch_count()
    ms_max, per_step, per_launch = timed_steps(b, args.steps, args.warmup, mark_mid=True)
    launches = L.ecg_launch_count() - launches0 - 2 * args.warmup if b.two_kernel_step else args.steps
    clocks = sampler.stop() if sampler else None
    handed_off = int(b._scratch[0].item()) if b.two_kernel_step else 0  # of the last step
    n_total = sum_over_ranks(n_main)
    value = n_total * args.steps / (ms_max * 1e-3)
    # health of the run: every board advanced every step
    stuck = int(((b.status & E.ST_NO_LEGAL) != 0).sum().item())
    bad = int(((b.status & ~E.ST_NO_LEGAL) != 0).sum().item())
    stats = E.dist.stats_dict(E.dist.reduce_stats(b.episode_stats()))
    mean_casc = float(b.cascades.float().mean().item())
    mean_reward = float(b.step_reward.float().mean().item())

    # ---- sustained: back-to-back steps for >= 5 s, clocks and power sampled over them
    sustained = None
    if not args.no_extra_legs and args.sustained_seconds > 0:
        k_sus = max(args.steps, int(args.sustained_seconds / (ms_max / args.steps * 1e-3)) + 1)
        sampler = ClockSampler(local) if rank == 0 else None
        ms_sus, _, _ = timed_steps(b, k_sus, 0)
        c = sampler.stop() if sampler else {}
        sustained = {"value": n_total * k_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": k_sus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / k_sus, "sm_mhz_median": c.get("sm_mhz"),
                     "sm_mhz_min": c.get("sm_mhz_min"), "power_w_median": c.get("power_w_median"),
                     "reasons": c.get("reasons"), "clock_samples": c.get("samples"),
                     "vs_burst": (n_total * k_sus / (ms_sus * 1e-3)) / value}

    # ---- e2e through the host-buffer API: the env contract (uint8 cells) and the 4-bit form
    def e2e_leg(fmt, expand=False):
        hs = E.HostStepper(env, chunks=args.e2e_chunks, obs_format=fmt, host_expand=expand)
        for _ in range(2):
            hs.step(hs.random_action())
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            a = hs.random_action()          # D2H: the action the host sends back in
            hs.step(a)                      # H2D: actions; D2H: obs, reward, done, won
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        d2h = hs.d2h_bytes + hs.action_d2h_bytes
        out = {"value": n_total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "obs_format": fmt, "d2h_bytes_per_board": d2h / n_main,
               "pcie_d2h_gb_s_per_gpu": d2h * args.e2e_steps / dt / 1e9,
               "pcie_d2h_gb_s_all_gpus": sum_over_ranks(d2h) * args.e2e_steps / dt / 1e9,
               "api": "HostStepper.random_action() + HostStepper.step(actions_host) -> (obs, reward, done, won) in "
                      f"pinned host memory, {args.e2e_chunks} chunks pipelined over CUDA streams; obs = " +
                      ("uint8 [N,9,9] cell values, int32 reward/actions (the env contract)" if fmt == "uint8" else
                       "uint8 [N,41] 4-bit cell codes (ecg_unpack_nibbles), int16 reward/actions")}
        if expand:
            out["transport"] = ("the observation crosses PCIe as 4-bit codes (41 B per board) and is widened to the uint8 "
                                f"contract form by {hs.expand_threads} host threads of libecg.so (ecg_host_expander_*) "
                                "inside the timed region, chunk by chunk as the copies land")
            out["host_expand_threads"] = hs.expand_threads
        elif fmt == "uint8":
            out["transport"] = "the uint8 observation itself crosses PCIe (81 B per board)"
        hs.close()
        del hs
        return out

    e2e = e2e_nib = e2e_direct = e2e_expand = None
    if not args.no_e2e:
        # the env-contract form twice: the uint8 cells copied as they are, and 4-bit codes widened on the host (same
        # bytes in the caller's buffer); "e2e" is the faster of the two, both are reported
        e2e_direct = e2e_leg("uint8")
        e2e_expand = e2e_leg("uint8", expand=True)
        e2e = dict(e2e_expand if e2e_expand["value"] > e2e_direct["value"] else e2e_direct)
        e2e_nib = e2e_leg("nibbles")

    env = b = None  # the main batch is done: free its 1.6 GB before the other legs allocate theirs
    torch.cuda.empty_cache()

    # ---- the other scaling mode on the same K steps
    twin = None
    if not args.no_extra_legs:
        if world == 1:
            twin = {"value": value, "ms_per_step": ms_max / args.steps, "boards_total": n_main,
                    "note": "N = 1: weak and strong are the same run"}
        else:
            n2, b02 = (args.boards, rank * args.boards) if strong_main else (count, first)
            env2 = make_env(n2, b02)
            ms2, _, _ = timed_steps(env2.board, args.steps, args.warmup)
            tot2 = sum_over_ranks(n2)
            twin = {"value": tot2 * args.steps / (ms2 * 1e-3), "ms_per_step": ms2 / args.steps, "boards_total": int(tot2),
                    "boards_per_gpu": n2}
            env2 = None
            torch.cuda.empty_cache()

    # ---- the reference's dynamics: every step restarts the MT19937 stream of cfg.seed (boardv2.py:46)
    replay = None
    if not args.no_extra_legs and args.replay_steps > 0:
        replay = {}
        seed = 20261019
        src = E.BatchedBoards(E.BoardConfig(seed=seed), n_main, 1 << 30, device=dev, key=KEY, board0=board0_main)
        for name, streams in (("shared_stream", 1), ("streams_4096", 4096)):
            kw = dict(seeds=[seed]) if streams == 1 else dict(
                seeds=[seed + i for i in range(streams)],
                stream_index=(torch.arange(n_main, device=dev, dtype=torch.int64) + board0_main).remainder(streams))
            rb = E.BatchedBoards(E.BoardConfig(seed=seed), n_main, 1 << 30, device=dev, refill="replay",
                                 stream_len=2048, env_goal=500, **kw)
            rb.boards.copy_(src.boards)  # distinct boards (Philox-drawn, no match on them), replayed refills and picks
            rb._mask_valid = False
            rb.packed_mask()
            ms_r, _, _ = timed_steps(rb, args.replay_steps, 3)
            overflow = int(((rb.status & E.ST_STREAM_OVERFLOW) != 0).sum().item())
            replay[name] = {"value": n_total * args.replay_steps / (ms_r * 1e-3), "unit": UNIT,
                            "ms_per_step": ms_r / args.replay_steps, "steps": args.replay_steps, "streams": streams,
                            "mean_cascades_per_step": float(rb.cascades.float().mean().item()),
                            "mean_reward_per_step": float(rb.step_reward.float().mean().item()),
                            "boards_with_stream_overflow": overflow}
            del rb
            torch.cuda.empty_cache()
        del src
        replay["kernel"] = ("two-kernel replay step: lane_kernel<Shape<9,9,3,false>, replay, step, FAST> reads its refill "
                            "tiles from per-stream tile tables (ecg_replay_tiles: the accepted values of numpy's masked "
                            "rejection over the raw MT19937 words, 16 per window), picks by masked rejection on the raw "
                            "words; rare cases go to lane_kernel<..., replay, step, exact>")
        replay["philox_for_comparison"] = {"mean_cascades_per_step": mean_casc, "mean_reward_per_step": mean_reward}

    # ---- BASELINE configs[3]: the board-shape sweep, each shape with its own roofline (per GPU, shards independent)
    sweep = None
    if not args.no_extra_legs and args.sweep_boards > 0:
        sweep = []
        peak_gbs, _ = measured_peak()
        for rows, types in ((6, 4), (9, 6), (12, 7), (16, 8)):
            cfg_s = E.BoardConfig(seed=5, rows=rows, columns=rows, types=types)
            sb = E.BatchedBoards(cfg_s, args.sweep_boards, 1 << 30, device=dev, key=99, board0=rank * args.sweep_boards)
            sb.packed_mask()
            ms_s, _, _ = timed_steps(sb, 12, 4)
            bytes_per_step = 2 * ((rows * rows + 1) // 2) + 17 + (cfg_s.action_space + 7) // 8
            rate = world * args.sweep_boards * 12 / (ms_s * 1e-3)
            sweep.append({"shape": f"{rows}x{rows}x{types}", "boards_per_gpu": args.sweep_boards, "value": rate,
                          "unit": UNIT, "ms_per_step": ms_s / 12, "algorithmic_bytes_per_step": bytes_per_step,
                          "roofline_frac": rate / world * bytes_per_step / 1e9 / peak_gbs,
                          "mean_cascades_per_step": float(sb.cascades.float().mean().item()),
                          "handed_off_to_exact_kernel": int(sb._scratch[0].item())})
            sb = None
            torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: mctslib standard MCTS, GPU-batched rollouts, reward sums reduced over NCCL
    mcts_leg = None
    if not args.no_extra_legs and args.mcts_sims > 0:
        state = E.BoardV2(20, E.BoardConfig(seed=7), device=dev)
        m = E.BatchedRolloutMCTS(state, 3, 2, False, leaves=args.mcts_leaves, key=1234)
        m()  # warm-up: 2 simulations (also re-roots the tree, like the reference's move loop)
        m._simulations = args.mcts_sims
        torch.cuda.synchronize(dev)
        barrier()
        steps0 = m.env_steps
        t0 = time.perf_counter()
        action, value_m, policies = m()
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        mcts_leg = {"leaves_per_simulation": args.mcts_leaves, "simulations": args.mcts_sims, "moves": 20,
                    "ms_per_simulation": dt / args.mcts_sims * 1e3, "leaves_per_s": args.mcts_leaves * args.mcts_sims / dt,
                    "rollout_env_steps_per_s": (m.env_steps - steps0) / dt, "root_children": len(policies),
                    "refill": "philox", "reduction": ("NCCL all-reduce of the reward sum per simulation (visit counts "
                                                      "are known on the host)") if world > 1 else "none (1 GPU)"}
        m = state = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_rate(12.0)
        try:  # the reference's own Python/NumPy path: measured where it can be imported, committed as a file
            with open(os.path.join(ROOT, "profiles", "python_reference_authoring_container.json")) as f:
                pr = json.load(f)
            cpu["python_reference"] = {
                "where": "NOT this box: the reference is pure Python under /root/reference of the authoring container "
                         "and cannot travel; scripts/time_python_reference.py ran it there",
                "cpu": pr["cpu"], "cores": pr["cores"],
                "single_process_env_steps_per_s": pr["single_process"]["env_steps_per_s"],
                "all_cores_reference_pool_env_steps_per_s": pr["all_cores_reference_pool"]["env_steps_per_s"]}
        except Exception:
            cpu["python_reference"] = "absent on this box (pure Python, cannot travel)"

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_launch_ms = sum(per_launch) / len(per_launch)
        avg_step_ms = sum(per_step) / len(per_step)
        finished = n_main - handed_off  # env-steps the dominant kernel completes per launch
        achieved = BYTES_PER_STEP * finished / (avg_launch_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        per_gpu = n_main
        scale_traffic = per_gpu / traffic["boards_per_launch"] if traffic and traffic.get("boards_per_launch") else 1.0
        cfg = workload(args.boards if not strong_main else n_main, world)
        cfg["scaling_mode"] = args.scaling
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e, "e2e_uint8_direct": e2e_direct, "e2e_uint8_host_expand": e2e_expand,
            "e2e_nibbles": e2e_nib, "gpu_launches": int(launches),
            ("weak" if strong_main else "strong"): twin, "sustained": sustained, "replay": replay,
            "shape_sweep": sweep, "mcts": mcts_leg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch") * scale_traffic,
                         "traffic_source": None if traffic is None else traffic.get("source"),
                         "kernel": "lane_kernel<Shape<9,9,3,false>, philox, step, FAST> (common-case kernel of the "
                                   "two-kernel step: persistent warp loop, free-running warps)",
                         "algorithmic_bytes_per_launch": BYTES_PER_STEP * finished, "avg_launch_ms": avg_launch_ms,
                         "env_steps_finished_per_launch": finished, "handed_off_to_exact_kernel": handed_off,
                         "exact_kernel_avg_ms": avg_step_ms - avg_launch_ms, "step_avg_ms": avg_step_ms,
                         "peak_source": peak_src,
                         "integer_pipe": None if traffic is None else {
                             "alu_pipe_pct_of_peak": traffic.get("alu_pipe_pct_of_peak"),
                             "issue_slots_pct_busy": traffic.get("issue_slots_pct_busy"),
                             "active_threads_per_instruction": traffic.get("active_threads_per_instruction"),
                             "source": "ncu capture of this kernel, profiles/step_kernel_traffic.json"},
                         "note": "integer-ALU-pipe bound, not HBM bound (DESIGN.md section 5): the ncu figures above "
                                 "are from the committed capture of this kernel, not from this run"},
            "cpu_baseline": cpu,
            "run": {"mean_cascades_per_step": mean_casc, "mean_reward_per_step": mean_reward,
                    "boards_without_legal_move": stuck, "boards_flagged": bad, "episode_stats": stats},
        }
        print(json.dumps(line))
    if world > 1:
        tdist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
