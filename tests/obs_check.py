"""Shared body of the observation parity check: drives an env (host-emulated or CUDA) and the oracle in lock
step and compares every decision row's legal mask and v4 observation rows 0..888."""
import numpy as np

import oracle_lib as O

EXP_ROWS = [131] + [132 + 195 * k + 192 + j for k in range(3) for j in range(3)]  # exp(-0.2 * age) planes


def check_obs_parity(make_env, fetch, n=8, max_cycles=400, seed0=777, min_rows=2000, sp=False, sp_tol=0.0, version=4):
    """make_env(nonces, keys) -> env ; fetch(env, first, prev_actions) -> (row_table, row_seat, masks, obs, actions)"""
    nonces = np.arange(seed0, seed0 + n, dtype=np.uint64)
    keys = np.full(n, 99, dtype=np.uint64)
    env = make_env(nonces, keys)
    L = O.lib()
    games = [L.orc_game_new(int(nonces[t]), int(keys[t]), 0, t) for t in range(n)]
    checked = 0
    exact = np.ones(889, dtype=bool)
    exact[EXP_ROWS] = False
    actions = None
    try:
        for cycle in range(max_cycles):
            rows_t, rows_s, masks, obs, actions = fetch(env, cycle == 0, actions)
            for t in range(n):
                assert L.orc_game_poll(games[t]) >= 0, O.err()
            chosen = {}
            for r in range(len(rows_t)):
                t, seat, kan = int(rows_t[r]), int(rows_s[r] & 3), bool(rows_s[r] & 4)
                ps = O.PlayerState(0, _ptr=L.orc_game_state(games[t], seat), _own=False)
                ref_obs, ref_mask = ps.encode_obs(version, kan, sp_mode=1 if sp else 0)
                assert (ref_mask == masks[r]).all(), (cycle, t, seat, kan)
                if version != 4:
                    # legacy layouts: cells the reference sets to exactly 0 or 1 must match exactly, the exp()-derived
                    # planes (RBF integer encodings, pond decay) to 1e-6
                    assert obs[r].shape == ref_obs.shape, (obs[r].shape, ref_obs.shape)
                    d = np.abs(obs[r] - ref_obs)
                    binary = (ref_obs == 0) | (ref_obs == 1)
                    bad = np.argwhere((d != 0) & binary)
                    assert len(bad) == 0, (version, cycle, t, seat, kan, bad[:8], obs[r][tuple(bad[0])], ref_obs[tuple(bad[0])])
                    assert d.max() <= 1e-6, (version, float(d.max()), np.unravel_index(np.argmax(d), d.shape))
                    checked += 1
                    chosen[(t, seat, kan)] = int(actions[r])
                    continue
                d = np.abs(obs[r][:889] - ref_obs[:889])
                bad = np.argwhere(d[exact] != 0)
                assert len(bad) == 0, (cycle, t, seat, kan, np.nonzero(exact)[0][bad[:8, 0]], bad[:8, 1],
                                       obs[r][:889][exact][tuple(bad[0])], ref_obs[:889][exact][tuple(bad[0])])
                assert d[~exact].max() <= 1e-6
                if sp:
                    dsp = np.abs(obs[r][889:] - ref_obs[889:])
                    if dsp.max() > sp_tol:
                        rr, cc = np.unravel_index(np.argmax(dsp), dsp.shape)
                        raise AssertionError(("sp block", cycle, t, seat, kan, 889 + int(rr), int(cc),
                                              float(obs[r][889 + rr, cc]), float(ref_obs[889 + rr, cc]), float(dsp.max())))
                else:
                    assert (obs[r][889:] == 0).all()
                checked += 1
                chosen[(t, seat, kan)] = int(actions[r])
            for (t, seat, kan), a in chosen.items():
                if kan:
                    continue
                ka = chosen.get((t, seat, True), -1)
                assert L.orc_game_set_action(games[t], seat, a, ka if a == 42 else -1) == 0, O.err()
            for t in range(n):
                L.orc_game_advance_step(games[t])
            if env.num_live() == 0:
                break
        assert checked >= min_rows, checked
    finally:
        for g in games:
            L.orc_game_free(g)
        env.close()
    return checked
