#!/usr/bin/env python
"""GPU debugging aid: run the obs parity check with SP and dump every differing cell of the first failing row."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mortal_b200, oracle_lib as O

n = 8
nonces = np.arange(777, 777 + n, dtype=np.uint64); keys = np.full(n, 99, dtype=np.uint64)
env = mortal_b200.BatchEnv(nonces, keys, enable_quick_eval=False)
L = O.lib()
games = [L.orc_game_new(int(nonces[t]), int(keys[t]), 0, t) for t in range(n)]
actions = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
bad_rows = 0
for cycle in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    env.step(None if cycle == 0 else actions)
    obs = env.encode_obs(); env.policy_test(1, actions)
    nr = env.num_rows()
    rt = env.row_table[:nr].cpu().numpy(); rs = env.row_seat[:nr].cpu().numpy()
    ob = obs[:nr].cpu().numpy(); acts = actions[:nr].cpu().numpy()
    for t in range(n): assert L.orc_game_poll(games[t]) >= 0
    chosen = {}
    for r in range(nr):
        t, seat, kan = int(rt[r]), int(rs[r] & 3), bool(rs[r] & 4)
        ps = O.PlayerState(0, _ptr=L.orc_game_state(games[t], seat), _own=False)
        ref, _ = ps.encode_obs(4, kan, sp_mode=1)
        d = np.argwhere(ob[r][889:] != ref[889:])
        if len(d):
            bad_rows += 1
            v = ps.view()
            print(f"cycle {cycle} table {t} seat {seat} kan {kan}: {len(d)} differing cells; shanten {v.shanten} rts {v.real_time_shanten} tiles_left {v.tiles_left} cans {v.cans:#x}")
            print("  tehai", [i for i in range(34) for _ in range(v.tehai[i])])
            for rr, cc in d[:40]:
                print(f"   row {889+rr} col {cc}: gpu {ob[r][889+rr, cc]} ref {ref[889+rr, cc]}")
        chosen[(t, seat, kan)] = int(acts[r])
    for (t, seat, kan), a in chosen.items():
        if kan: continue
        ka = chosen.get((t, seat, True), -1)
        assert L.orc_game_set_action(games[t], seat, a, ka if a == 42 else -1) == 0
    for t in range(n): L.orc_game_advance_step(games[t])
print("bad rows", bad_rows, "overflows", env.sp_overflows())
