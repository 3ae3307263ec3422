"""Run under torchrun (one rank per GPU): BASELINE configs[4] in miniature. Every rank plays its shard of short hanchans to
the end on its own GPU, the real returns {scores i32[4], rank u8[4]} are all-gathered with mortal_b200.dist.gather_returns
(NCCL), and every rank checks the WHOLE gathered array against the oracle (test infrastructure). Prints RETURNS_OK on rank 0."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds-per-rank", type=int, default=64)
    ap.add_argument("--backend", default="nccl")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    if a.backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(a.backend)
    import mortal_b200
    from mortal_b200 import dist as mdist

    seed_start = (10000, 0x2000)
    nonces, keys = mdist.shard_seeds(seed_start, a.seeds_per_rank, rank)
    env = mortal_b200.BatchEnv(nonces, keys, device=local)
    res = env.run_test_policy(kind=1)
    env.close()
    assert (res["err"] == 0).all() and (res["done"] == 1).all()
    scores, ranks = mdist.gather_returns(res["scores"], res["ranks"], device=torch.device("cuda", local))
    n = len(nonces)
    assert scores.shape == (world * n, 4) and (scores[rank * n:(rank + 1) * n] == res["scores"]).all()
    import oracle_lib as O

    all_nonces = np.concatenate([mdist.shard_seeds(seed_start, a.seeds_per_rank, r)[0] for r in range(world)])
    all_keys = np.full(len(all_nonces), seed_start[1], dtype=np.uint64)
    # table ids are per-rank local (the counter-based test policy hashes the table index)
    tids = np.tile(np.arange(n, dtype=np.int32), world)
    ref = O.run_batch(all_nonces, all_keys, policy_kind=1, n_threads=min(32, os.cpu_count() or 1), table_ids=tids)
    assert (ref["scores"] == scores).all() and (ref["ranks"] == ranks).all(), f"rank {rank}: gathered returns differ from the oracle"
    assert (scores.sum(1) == 100000).all()
    ok = torch.ones(1, device=torch.device("cuda", local) if a.backend == "nccl" else "cpu")
    dist.all_reduce(ok)
    if rank == 0:
        print(f"RETURNS_OK world={world} tables={world * n}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
