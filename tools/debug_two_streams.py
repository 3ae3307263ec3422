#!/usr/bin/env python
"""Experiment: the env-only loop as K independent table groups on K CUDA streams (same 4096 tables in total).
Prints ms per whole-batch cycle for K = 1, 2, 4. Answers whether kernel-level concurrency hides the latency-bound
kernels (k_step, k_encode_features, small single-player levels) behind the other group's work."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mortal_b200

N, FF, CYC = 4096, 300, 20
for K in (1, 2, 4):
    per = N // K
    envs, acts, streams = [], [], []
    for g in range(K):
        lo = 10000 + g * (per // 4)
        nonces = np.repeat(np.arange(lo, lo + per // 4, dtype=np.uint64), 4)
        env = mortal_b200.BatchEnv(nonces, np.full(per, 0x2000, dtype=np.uint64))
        envs.append(env)
        acts.append(torch.zeros(env.row_cap, dtype=torch.int64, device=env.device))
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    for env, a, s in zip(envs, acts, streams):
        with torch.cuda.stream(s):
            env.step(None)
            env.policy_test(2, a)
            for _ in range(FF):
                env.step(a)
                env.policy_test(2, a)
    torch.cuda.synchronize()

    def cycle():
        for env, a, s in zip(envs, acts, streams):
            with torch.cuda.stream(s):
                env.step(a)
                env.encode_obs()
                env.policy_test(2, a)

    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in streams:
        s.wait_event(t0)
    for _ in range(CYC):
        cycle()
    for s in streams:
        e = torch.cuda.Event()
        e.record(s)
        torch.cuda.current_stream().wait_event(e)
    t1.record()
    torch.cuda.synchronize()
    ov = sum(env.sp_overflows() for env in envs)
    print(f"K={K}: {t0.elapsed_time(t1) / CYC:.3f} ms per {N}-table cycle, sp_overflows {ov}")
    del envs, acts
