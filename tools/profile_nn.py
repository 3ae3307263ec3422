#!/usr/bin/env python
"""Times the policy-net forward alone (192x40 brain + DQN, bf16 fast path) at a given batch, with / without cudnn.benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mortal_b200.engine import DeviceEngine
from mortal_b200.model import DQN, Brain

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4022)
ap.add_argument("--benchmark", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
torch.backends.cudnn.benchmark = bool(args.benchmark)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
eng = DeviceEngine(Brain(conv_channels=192, num_blocks=40, version=4), DQN(version=4), device=dev)
obs = torch.rand((args.batch, 1012, 34), device=dev)
masks = torch.rand((args.batch, 46), device=dev) > 0.5
masks[:, 45] = True
for _ in range(3):
    eng.react_device(obs, masks)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.iters):
    eng.react_device(obs, masks)
t1.record()
torch.cuda.synchronize()
print(f"batch {args.batch} cudnn.benchmark={args.benchmark}: {t0.elapsed_time(t1) / args.iters:.2f} ms / forward")
