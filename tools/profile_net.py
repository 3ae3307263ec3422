#!/usr/bin/env python
"""The policy-net forward alone (192x40 brain, bf16 fast path) at bench.py's row count, for profilers:
`ncu --metrics gpu__time_duration.sum ... python tools/profile_net.py` then tools/launch_hist.py on the CSV."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mortal_b200.engine import DeviceEngine
from mortal_b200.model import DQN, Brain

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
torch.manual_seed(0)
dev = torch.device("cuda", 0)
eng = DeviceEngine(Brain(conv_channels=192, num_blocks=40, version=4), DQN(version=4), device=dev, enable_amp=True, name="m")
obs = (torch.rand((rows, 1012, 34), device=dev) < 0.03).float()
masks = torch.ones((rows, 46), dtype=torch.bool, device=dev)
for _ in range(3):
    eng.react_device(obs, masks)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
torch.cuda.nvtx.range_push("forward")
eng.react_device(obs, masks)
torch.cuda.nvtx.range_pop()
t1.record()
torch.cuda.synchronize()
print(f"forward {rows} rows: {t0.elapsed_time(t1):.3f} ms (eager launches)")
