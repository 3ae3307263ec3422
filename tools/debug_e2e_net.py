#!/usr/bin/env python
"""Per-cycle wall time of the arena driven by DeviceEngine.react_batch over host buffers (bench.py's e2e_with_net path)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mortal_b200.libriichi as lr
from bench import HostNetEngine
from mortal_b200.engine import DeviceEngine
from mortal_b200.model import DQN, Brain

lr.install()
from libriichi.arena import OneVsThree

torch.manual_seed(0)
dev = torch.device("cuda", 0)
eng = DeviceEngine(Brain(conv_channels=192, num_blocks=40, version=4), DQN(version=4), device=dev, enable_amp=True, enable_quick_eval=True, name="m")
arena = OneVsThree(disable_progress_bar=True)
arena.fast_forward_steps = 300
arena.max_cycles = 16
last = [time.perf_counter()]


def hook(c, state):
    torch.cuda.synchronize()
    now = time.perf_counter()
    print(f"cycle {c}: {(now - last[0]) * 1e3:.1f} ms", flush=True)
    last[0] = now


arena.cycle_hook = hook
agent = HostNetEngine(eng)
arena.py_vs_py(agent, agent, (10000, 0x2000), 1024)
