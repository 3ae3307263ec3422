#!/usr/bin/env python
"""GameplayLoader throughput: device log replay vs the oracle restatement on the host cores (moves/s, v4 obs + SP block).
Logs are produced on the spot by the arena (greedy-ish random engine), so the tool needs no dataset."""
import argparse
import concurrent.futures as cf
import gzip
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from mortal_b200 import dataset_codec as DC
from mortal_b200.libriichi.arena import OneVsThree
from mortal_b200.libriichi.dataset import GameplayLoader

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=16)
ap.add_argument("--cpu-jobs", type=int, default=32, help="(log, player) jobs timed on the CPU oracle")
args = ap.parse_args()


class Eng:
    engine_type, version, is_oracle, enable_quick_eval, enable_rule_based_agari_guard, name = "mortal", 4, False, True, False, "e"

    def react_device(self, obs, masks):
        q = torch.rand(masks.shape, device=masks.device)
        q[:, :34] += 2.0 * obs[:, 876, :] + obs[:, 875, :]
        q = q.masked_fill(~masks, -1.0)
        return q.argmax(-1), q


with tempfile.TemporaryDirectory() as d:
    OneVsThree(disable_progress_bar=True, log_dir=d).py_vs_py(Eng(), Eng(), (42000, 5), args.seeds)
    files = sorted(os.path.join(d, f) for f in os.listdir(d))
    loader = GameplayLoader(4, oracle=False)
    loader.load_gz_log_files(files[:4])  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loaded = loader.load_gz_log_files(files)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    moves = sum(len(gp.take_actions()) for g in loaded for gp in g)
    print(f"device: {len(files)} logs x 4 players, {moves} moves in {dt:.2f} s = {moves / dt:.0f} moves/s (incl. gz + json parsing on host)")
    import oracle_lib as O

    O.lib()
    jobs = [(f, p) for f in files for p in range(4)][: args.cpu_jobs]
    texts = {f: DC.parse_log(gzip.open(f, "rt").read()) for f, _ in jobs}

    def one(job):
        f, p = job
        return len(O.gameplay_load(texts[f], p, version=4, sp_mode=1)["actions"])

    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(cores) as ex:  # the ctypes call releases the GIL
        cpu_moves = sum(ex.map(one, jobs))
    dt = time.perf_counter() - t0
    print(f"oracle: {len(jobs)} jobs, {cpu_moves} moves in {dt:.2f} s = {cpu_moves / dt:.0f} moves/s on {cores} host threads")
