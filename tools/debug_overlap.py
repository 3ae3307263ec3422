#!/usr/bin/env python
"""Does a D2H copy overlap the single-player kernels on this box? Times the pieces of mjx_env_encode_obs_host separately."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mortal_b200

n = 4096
env = mortal_b200.BatchEnv(np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4), np.full(n, 0x2000, dtype=np.uint64))
acts = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
env.step(None); env.policy_test(1, acts)
for _ in range(300):
    env.step(acts); env.policy_test(1, acts)
obs = env.obs_buffer()
h_obs = torch.empty((env.row_cap, 1012, 34), dtype=torch.float32).pin_memory()
h_masks = torch.empty((env.row_cap, 46), dtype=torch.bool).pin_memory()
other = torch.empty((4096, 1012, 34), dtype=torch.float32, device=env.device)
h_other = torch.empty((4096, 1012, 34), dtype=torch.float32).pin_memory()
side = torch.cuda.Stream()


def timed(fn, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def dev_only():
    env.step(acts); env.encode_obs(obs); env.policy_test(1, acts)


def copy_only():
    h_other.copy_(other, non_blocking=True)


def both_streams():  # device encode on the current stream, an unrelated D2H copy on a side stream, concurrently
    with torch.cuda.stream(side):
        h_other.copy_(other, non_blocking=True)
    env.step(acts); env.encode_obs(obs); env.policy_test(1, acts)


def host_path():
    env.step(acts); env.encode_obs_host(h_obs, h_masks); env.policy_test(1, acts)


for name, fn in (("device step+encode", dev_only), ("plain D2H of 4096 obs", copy_only), ("both concurrently", both_streams),
                 ("encode_obs_host", host_path)):
    timed(fn, 2)
    print(f"{name:28s} {timed(fn):7.2f} ms")
