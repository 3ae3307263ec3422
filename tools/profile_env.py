#!/usr/bin/env python
"""Env-only loop for profilers (no network): `ncu ... python tools/profile_env.py --cycles 12 [--no-sp]`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mortal_b200

ap = argparse.ArgumentParser()
ap.add_argument("--tables", type=int, default=4096)
ap.add_argument("--cycles", type=int, default=12)
ap.add_argument("--no-sp", action="store_true")
ap.add_argument("--skip", type=int, default=300, help="fast-forward batch steps (no encode) before the profiled cycles")
args = ap.parse_args()
n = args.tables
nonces = np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4)
keys = np.full(n, 0x2000, dtype=np.uint64)
env = mortal_b200.BatchEnv(nonces, keys)
env.set_sp(not args.no_sp)
actions = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env.step(None)
env.policy_test(2, actions)
for _ in range(args.skip):
    env.step(actions)
    env.policy_test(2, actions)
for i in range(args.cycles):
    if i == 3:
        t0.record()
    env.step(actions)
    env.encode_obs()
    env.policy_test(2, actions)
t1.record()
torch.cuda.synchronize()
print(f"{args.cycles - 3} cycles, {t0.elapsed_time(t1) / max(args.cycles - 3, 1):.3f} ms/cycle, rows last {env.num_rows()}, "
      f"sp_overflows {env.sp_overflows()}, sp states/edges/slots {env.sp_stats() if not args.no_sp else None}")
