#!/usr/bin/env python
"""Benchmark: table-steps/sec of batched riichi self-play (BASELINE.json metric).

A "step" = one iteration of libriichi's BatchGame::run loop (arena/game.rs:286-304) over the whole batch:
commit the previous decisions, poll every live table to its next decision point, encode one v4
observation per decision row, run the policy. One table-step = that iteration for one live table.

Default arm (this repo): 4096 tables per GPU (BASELINE configs[1]), random-init Mortal brain
(192 channels x 40 blocks, bf16 autocast, greedy), everything resident in HBM. JSON also carries
  env_only      the same loop with the counter-based test policy instead of the network
  roofline      achieved HBM GB/s of the HBM-bound env kernel (k_encode_store; .pair = with k_encode_features) from CUDA events
  e2e           the loop through the C ABI with HOST buffers (obs D2H, actions H2D every step)
  cpu_baseline  the CPU oracle on this box's host cores, bounded sample (rank 0, N=1 only)
`--impl reference` times libriichi's own CPU path restated by the oracle (oracle/, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TABLES = 4096
SEED_START = (10000, 0x2000)  # mortal/player.py:67
OBS_BYTES = 1012 * 34 * 4
MASK_BYTES = 46
STATE_BYTES = 1952  # sizeof(TableState) read per encoded row


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def seeds_for_rank(rank: int, n_tables: int):
    import numpy as np

    count = n_tables // 4
    start = SEED_START[0] + count * rank  # rank r takes seed_start + 1024 r
    nonces = np.repeat(np.arange(start, start + count, dtype=np.uint64), 4)
    keys = np.full(n_tables, SEED_START[1], dtype=np.uint64)
    return nonces, keys


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """libriichi's CPU path (oracle restatement; the Rust crate cannot be built here): poll/commit loop +
    one v4 obs encode (incl. the single-player tables) per decision row, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O

    cores = host_cores()
    n_tables = args.ref_tables
    steps_per_table = args.ref_steps_per_table
    nonces, keys = seeds_for_rank(0, N_TABLES)
    nonces, keys = nonces[:n_tables], keys[:n_tables]

    def one():
        r = O.run_batch(nonces, keys, shuffle_kind=0, policy_kind=1, quick_eval=True, encode_obs=4, sp_mode=1,
                        n_threads=cores, max_steps=args.skip + steps_per_table, encode_from_step=args.skip)
        return r["table_steps"], r["seconds"], r["obs_rows"]

    for _ in range(args.warmup):
        one()
    tot_steps, tot_sec = 0, 0.0
    for _ in range(args.steps):
        s, t, _ = one()
        tot_steps += s
        tot_sec += t
    value = tot_steps / tot_sec
    sample = f"{n_tables} tables x table-steps {args.skip}..{args.skip + steps_per_table} each per step (seeds {SEED_START[0]}.., greedy test policy, v4 obs + SP encode per decision)"
    line = {
        "impl": "reference", "metric": "table-steps/sec batched self-play", "value": value, "unit": "table-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 (+f32 SP)", "data": "synthetic",
        "config": {"workload": "BatchGame 4096 tables self-play step loop (configs[1]), CPU arena", "tables_per_gpu": N_TABLES,
                   "obs_version": 4, "policy": "counter-based greedy test policy (no network on the CPU arm)"},
        "cpu_baseline": {"value": value, "unit": "table-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "table-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mortal_b200
    from mortal_b200.engine import DeviceEngine
    from mortal_b200.model import DQN, Brain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mortal_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nonces, keys = seeds_for_rank(rank, N_TABLES)
    torch.manual_seed(0)
    engine = DeviceEngine(Brain(conv_channels=192, num_blocks=40, version=4), DQN(version=4), device=dev,
                          enable_amp=True, enable_quick_eval=True)

    def fresh_env():
        """A new batch, fast-forwarded (untimed, greedy test policy, no encode) by --skip batch steps so that the timed
        steps see the steady-state mix of early/late kyoku positions instead of 4096 freshly dealt hands."""
        env = mortal_b200.BatchEnv(nonces, keys, obs_version=4, shuffle_kind=0, enable_quick_eval=True, device=local_rank)
        actions = torch.zeros(env.row_cap, dtype=torch.int64, device=dev)
        env.step(None)
        env.policy_test(1, actions)
        for _ in range(args.skip):
            env.step(actions)
            env.policy_test(1, actions)
        return env, actions

    W, K = args.warmup, args.steps
    stats = {}

    # -------- loop A: with the network (the BASELINE config), HBM resident
    def loop(env_actions, policy, n_warm, n_timed, time_encode=False, split_events=None):
        env, actions = env_actions
        obs = env.obs_buffer()
        enc_events = []

        def cycle(timed):
            if timed and split_events is not None:  # where the step goes: env kernels vs policy (events only, no sync)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                env.step(actions)
                env.encode_obs(obs)
                ev[1].record()
                r = policy(env, obs, actions)
                ev[2].record()
                split_events.append(ev)
                return r
            env.step(actions)
            if timed and time_encode:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                env.encode_obs(obs)
                e1.record()
                enc_events.append((e0, e1))
            else:
                env.encode_obs(obs)
            return policy(env, obs, actions)

        for _ in range(n_warm):
            cycle(False)
        barrier()
        steps0, rows, l0 = env.total_steps(), 0, env.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_timed):
            rows += cycle(True)
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        steps = env.total_steps() - steps0
        enc_ms = sum(a.elapsed_time(b) for a, b in enc_events)
        return dict(ms=ms, table_steps=steps, rows=rows, enc_ms=enc_ms, n=n_timed, launches=env.launch_count() - l0)

    def nn_policy(env, obs, actions):
        nr = env.num_rows()  # the only host sync of the cycle: the batch size the network runs at
        if nr:
            a, _ = engine.react_static(obs, env.masks, nr)  # CUDA-graph replay over the env's persistent buffers
            actions[:nr] = a
        return nr

    def test_policy(env, obs, actions):
        env.policy_test(1, actions)
        return 0

    sampler = ClockSampler(local_rank)
    sampler.start()
    ea = fresh_env()
    a_split = []
    a = loop(ea, nn_policy, W, K, split_events=a_split)
    a_env_ms = sum(e[0].elapsed_time(e[1]) for e in a_split) / K
    a_nn_ms = sum(e[1].elapsed_time(e[2]) for e in a_split) / K
    ea[0].close()
    clocks = sampler.stop()

    # -------- loop B: env only (test policy on device, no host sync)
    ea = fresh_env()
    b = loop(ea, test_policy, W, K)
    sp_overflows = ea[0].sp_overflows()
    sp_states, sp_edges, _ = ea[0].sp_stats()  # size of the last step's single-player DP
    ea[0].close()
    # -------- loop B2: the HBM-bound encode kernel alone (single-player block off), timed with CUDA events
    ea = fresh_env()
    ea[0].set_sp(False)
    b2 = loop(ea, test_policy, W, K, time_encode=True)
    ea[0].close()
    # rows per launch for the roofline: the same deterministic K cycles again, reading the row count each step
    # and the two encoder kernels timed separately (events inside libmjx on the launch stream; a sync per step, so this
    # pass is not the one `env_only` is quoted from)
    env, actions = fresh_env()
    env.set_sp(False)
    env.set_encode_timing(True)
    obs_t = env.obs_buffer()
    b_rows, feat_ms, store_ms = 0, 0.0, 0.0
    for i in range(W + K):
        env.step(actions)
        env.encode_obs(obs_t)
        env.policy_test(1, actions)
        if i >= W:
            b_rows += env.num_rows()
            f_ms, s_ms = env.last_encode_ms()
            feat_ms += f_ms
            store_ms += s_ms
    env.close()

    # -------- BASELINE configs[3]: encode_obs throughput at 65536 decision rows per launch (rank 0, N=1 only: it is a
    # kernel measurement, not part of the step). 65536 tables, one row per table-step on average, SP block off.
    enc64k = None
    if world == 1 and not args.no_encode_64k:
        n64 = 65536
        n_nonce = np.repeat(np.arange(SEED_START[0], SEED_START[0] + n64 // 4, dtype=np.uint64), 4)
        env = mortal_b200.BatchEnv(n_nonce, np.full(n64, SEED_START[1], dtype=np.uint64), obs_version=4, device=local_rank)
        env.set_sp(False)
        acts = torch.zeros(env.row_cap, dtype=torch.int64, device=dev)
        env.step(None)
        env.policy_test(1, acts)
        for _ in range(60):
            env.step(acts)
            env.policy_test(1, acts)
        obs64 = env.obs_buffer()
        rows64, ms64 = 0, 0.0
        for i in range(3 + 5):
            env.step(acts)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.encode_obs(obs64)
            e1.record()
            env.policy_test(1, acts)
            nr64 = env.num_rows()
            if i >= 3:
                rows64 += nr64
                ms64 += e0.elapsed_time(e1)
        env.close()
        del obs64
        torch.cuda.empty_cache()
        gbs64 = rows64 * (OBS_BYTES + MASK_BYTES + STATE_BYTES) / (ms64 * 1e-3) / 1e9
        enc64k = {"rows_per_launch": rows64 / 5, "ms_per_launch": ms64 / 5, "achieved": gbs64, "unit": "GB/s"}

    # -------- loop C: e2e through the C ABI with HOST buffers (pinned): obs/masks D2H, actions H2D each step
    env, d_actions = fresh_env()
    h_obs = torch.empty((env.row_cap, 1012, 34), dtype=torch.float32, pin_memory=True)
    h_masks = torch.empty((env.row_cap, 46), dtype=torch.bool, pin_memory=True)
    h_actions = torch.zeros(env.row_cap, dtype=torch.int64).pin_memory()
    h_actions.copy_(d_actions)
    # host policy standing in for engine.react_batch on HOST tensors: greedy on the observation it was handed
    # (agari > riichi > shanten-lowering discard > shanten-keeping discard > pass > calls), so that the hands keep
    # developing the way they do under a real policy and the single-player block stays as expensive as in self-play
    prio = torch.zeros(46)
    prio[43], prio[37], prio[44], prio[45] = 100.0, 50.0, 10.0, 0.5
    prio[38:43] = 0.25
    prio += torch.arange(46, dtype=torch.float32) * 1e-4
    aka_base = torch.tensor([4, 13, 22])

    def host_policy(nr):
        m = h_masks[:nr]
        score = prio.repeat(nr, 1)
        disc = 2.0 * h_obs[:nr, 876, :] + h_obs[:nr, 875, :] + 1.0  # v4 rows 875/876: keep / next shanten discards
        score[:, :34] += disc
        score[:, 34:37] += disc[:, aka_base] - 0.5
        score[~m] = -1.0
        return score.argmax(-1)

    e2e_stage = {"step_encode_d2h": 0.0, "host_policy": 0.0}

    def e2e_cycle():
        ta = time.perf_counter()
        d_actions.copy_(h_actions, non_blocking=True)  # H2D: the step's inputs
        env.step(d_actions)
        nr = env.encode_obs_host(h_obs, h_masks)  # D2H: the step's result, as react_batch receives it (blocking)
        tb = time.perf_counter()
        if nr:
            h_actions[:nr] = host_policy(nr)
        e2e_stage["step_encode_d2h"] += tb - ta
        e2e_stage["host_policy"] += time.perf_counter() - tb
        return nr

    for _ in range(W):
        e2e_cycle()
    barrier()
    e2e_stage["step_encode_d2h"] = e2e_stage["host_policy"] = 0.0
    s0 = env.total_steps()
    w0 = time.perf_counter()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_rows = 0
    for _ in range(K):
        e2e_rows += e2e_cycle()
    t1.record()
    barrier()
    e2e_ms = max(t0.elapsed_time(t1), (time.perf_counter() - w0) * 1000.0)
    e2e_steps = env.total_steps() - s0
    # what the link gives for the same bytes: one plain pinned D2H copy of the obs buffer (context for e2e, not a claim)
    nprobe = max(1, e2e_rows // K)
    dsrc = env.obs_buffer()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h_obs[:nprobe].copy_(dsrc[:nprobe], non_blocking=True)
    torch.cuda.synchronize()
    p0.record()
    h_obs[:nprobe].copy_(dsrc[:nprobe], non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    pcie_gbs = nprobe * OBS_BYTES / (p0.elapsed_time(p1) * 1e-3) / 1e9
    env.close()

    # -------- reduce over ranks: max time, sum of units
    def reduce(ms, units):
        if world == 1:
            return ms, units
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        u = torch.tensor([units], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        return float(t.item()), float(u.item())

    a_ms, a_units = reduce(a["ms"], a["table_steps"])
    b_ms, b_units = reduce(b["ms"], b["table_steps"])
    c_ms, c_units = reduce(e2e_ms, e2e_steps)

    # -------- the one collective of the path: all-gather of end-of-hanchan returns
    gather_us = None
    if world > 1:
        ret = torch.zeros((N_TABLES, 5), dtype=torch.int32, device=dev)  # scores[4] + packed ranks
        out = torch.empty((world * N_TABLES, 5), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(out, ret)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(out, ret)
        g1.record()
        torch.cuda.synchronize()
        gather_us = g0.elapsed_time(g1) * 1000.0

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        rows_per_launch = b_rows / K
        bytes_per_launch = rows_per_launch * (OBS_BYTES + MASK_BYTES + STATE_BYTES)
        enc_ms_per_launch = b2["enc_ms"] / K
        achieved = bytes_per_launch / (enc_ms_per_launch * 1e-3) / 1e9 if enc_ms_per_launch > 0 else 0.0
        line = {
            "metric": "table-steps/sec batched self-play", "value": a_units / (a_ms * 1e-3), "unit": "table-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": a_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/i32 env + bf16 policy net", "data": "synthetic",
            "config": {"workload": "BatchGame 4096 tables/GPU, random-init Mortal brain (192ch x 40 blocks, v4 obs), self-play step loop (BASELINE configs[1])",
                       "tables_per_gpu": N_TABLES, "global_tables": N_TABLES * world, "obs_version": 4,
                       "seed_start": list(SEED_START), "fast_forward_steps": args.skip, "parallelism": f"tables sharded dp{world}, no data-path collective",
                       "l2": "per-step obs output (~0.7 GB) exceeds the 126 MB L2, no explicit flush",
                       "sp_block": "rows 889-1011 (single-player tables) computed on device by the k_sp_* kernels",
                       "sp_arena_overflows": sp_overflows},
            # the timed step split with CUDA events (rank 0): env kernels (k_step + encode + single-player block, whose cost
            # depends on the positions the policy steers the tables into) and the policy network incl. the row-count sync
            "step_breakdown_ms": {"env": a_env_ms, "policy_net": a_nn_ms},
            "env_only": {"value": b_units / (b_ms * 1e-3), "unit": "table-steps/s", "ms_per_step": b_ms / K,
                         "policy": "counter-based test policy kernel, no host sync",
                         "without_sp_block": {"value": b2["table_steps"] / (b2["ms"] * 1e-3), "ms_per_step": b2["ms"] / K}},
            # the HBM-bound kernel of the path: k_encode_store materialises and stores every observation of the step (the
            # algorithmic bytes); k_encode_features, which derives the 11 KB/row compact form it reads, is latency-bound and
            # is reported beside it (`pair` = both kernels together, the figure earlier rounds quoted)
            "roofline": {"kernel": "k_encode_store", "bound": "hbm", "achieved": bytes_per_launch / (store_ms / K * 1e-3) / 1e9,
                         "peak": peak_gbs, "unit": "GB/s", "frac": bytes_per_launch / (store_ms / K * 1e-3) / 1e9 / peak_gbs,
                         # dram__bytes_read + dram__bytes_write of one `ncu --set full` capture at this workload
                         # (profiles/r01_ncu_k_encode_store.md): 44.2 MB read + 497.6 MB written per launch
                         "traffic": 541.8e6, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                         "ms_per_launch": store_ms / K, "rows_per_launch": rows_per_launch,
                         "k_encode_features_ms": feat_ms / K,
                         "pair": {"ms_per_launch": enc_ms_per_launch, "achieved": achieved, "frac": achieved / peak_gbs if peak_gbs else None,
                                  "traffic": 551.6e6}},
            # the single-player block is a latency-bound graph DP (hash interning + value propagation over an arena far larger
            # than L2); it has no meaningful HBM roofline, so it is reported as states/s. ms = env_only minus the same loop
            # with the block switched off.
            "sp_block": {"ms_per_step": (b["ms"] - b2["ms"]) / K, "states_last_step": sp_states, "edges_last_step": sp_edges,
                         "states_per_s": sp_states / max((b["ms"] - b2["ms"]) / K * 1e-3, 1e-9), "share_of_env_step": 1.0 - b2["ms"] / b["ms"]},
            # BASELINE configs[3] (encode_obs throughput, 65536 states -> obs tensor): the same two kernels at 16x the rows
            "encode_65536": (dict(enc64k, frac=enc64k["achieved"] / peak_gbs) if enc64k else None),
            "e2e": {"value": c_units / (c_ms * 1e-3), "unit": "table-steps/s",
                    "h2d_bytes_per_step": 8 * N_TABLES * 3,
                    "d2h_bytes_per_step": int(e2e_rows / K * (OBS_BYTES + MASK_BYTES)),
                    "path": "mjx_env_encode_obs_host with pinned host buffers: actions H2D, obs+masks D2H every step "
                            "(the single-player block runs in 4 row groups; finished groups drain through the copy engine meanwhile), "
                            "greedy host-side policy reading the host obs",
                    "plain_d2h_copy_gbs": pcie_gbs,
                    "stages_ms_per_step": {k: 1000.0 * v / K for k, v in e2e_stage.items()}},
            # this library's kernels in the timed region: env kernels counted by libmjx, plus the fused policy-net kernels
            # (4 per residual block + 1, csrc/mjx_nn.cuh) that each CUDA-graph replay of the forward contains
            "gpu_launches": a["launches"] + K * (4 * 40 + 1), "gpu_launches_env": a["launches"], "clocks": clocks,
            "collective": {"all_gather_us": gather_us, "bytes_per_table": 20},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    cores = host_cores()
    nonces, keys = seeds_for_rank(0, N_TABLES)
    n = args.ref_tables
    r = O.run_batch(nonces[:n], keys[:n], shuffle_kind=0, policy_kind=1, quick_eval=True, encode_obs=4, sp_mode=1,
                    n_threads=cores, max_steps=args.skip + args.ref_steps_per_table, encode_from_step=args.skip)
    return {"value": r["table_steps"] / r["seconds"], "unit": "table-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} tables x table-steps {args.skip}..{args.skip + args.ref_steps_per_table} (same fast-forward as the GPU arm), "
                      f"v4 obs + SP encode per decision, {r['seconds']:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--skip", type=int, default=300, help="untimed fast-forward batch steps before warm-up (both arms)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-tables", type=int, default=256)
    ap.add_argument("--ref-steps-per-table", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encode-64k", action="store_true", help="skip the BASELINE configs[3] encode measurement (27 GB obs buffer)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
