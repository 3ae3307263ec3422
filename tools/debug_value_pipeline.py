#!/usr/bin/env python
"""`value` workload (DeviceEngine, 192x40 net, 4096 tables) through the arena: one batch vs two half-batches on two streams."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mortal_b200.libriichi as lr
from mortal_b200.engine import DeviceEngine
from mortal_b200.model import DQN, Brain

lr.install()
from libriichi.arena import OneVsThree

torch.manual_seed(0)
dev = torch.device("cuda", 0)
eng = DeviceEngine(Brain(conv_channels=192, num_blocks=40, version=4), DQN(version=4), device=dev, enable_amp=True, enable_quick_eval=True, name="m")
for flag in (False, True, False, True):
    arena = OneVsThree(disable_progress_bar=True)
    arena.fast_forward_steps = 300
    arena.pipeline_device_engines = flag
    arena.max_cycles = 36
    marks = {}

    def hook(c, state):
        if c in (10, 35):
            torch.cuda.synchronize()
            marks[c] = (time.perf_counter(), state.total_steps())

    arena.cycle_hook = hook
    arena.py_vs_py(eng, eng, (10000, 0x2000), 1024)
    (t0, s0), (t1, s1) = marks[10], marks[35]
    print(f"pipeline_device_engines={flag}: {(s1 - s0) / (t1 - t0) / 1e3:.1f} K table-steps/s, {(t1 - t0) / 25 * 1e3:.2f} ms/cycle", flush=True)
