"""Rule-logic parity of the product's step sources (host-emulated, single lane) against the oracle.

This is the GPU-less half of the parity story: the same mjx_step.cuh that the CUDA kernel compiles is
built with -DMJX_HOST_EMUL and driven by the shared counter-based test policies; every decision row
(table, step, seat, legal-mask bits, action) and every final score / rank / step count must be equal.
The `-m gpu` tests repeat this through the real kernels and the C ABI.
"""
import json

import numpy as np
import pytest

import emul_lib as E
import oracle_lib as O


def sort_trace(t):
    # (table, step, seat, kan) is a unique key
    order = np.lexsort((t[:, 4], t[:, 2], t[:, 1], t[:, 0]))
    return t[order]


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if (a[i] != b[i]).any():
            return i, a[i], b[i]
    return n, None, None


@pytest.mark.parametrize("policy_kind", [1, 0, 2])
@pytest.mark.parametrize("quick_eval", [True, False])
@pytest.mark.parametrize("shuffle_kind", [0, 1])
def test_selfplay_trace_parity(policy_kind, quick_eval, shuffle_kind):
    n = 48
    nonces = np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4)  # OneVsThree seed layout
    keys = np.full(n, 0x2000, dtype=np.uint64)
    cap = 1 << 17
    ro = O.run_batch(nonces, keys, shuffle_kind=shuffle_kind, policy_kind=policy_kind, quick_eval=quick_eval,
                     trace_cap=cap)
    re = E.run(nonces, keys, shuffle_kind=shuffle_kind, policy_kind=policy_kind, quick_eval=quick_eval, trace_cap=cap)
    assert (re["errs"] == 0).all(), re["errs"]
    to, te = sort_trace(ro["trace"]), sort_trace(re["trace"])
    i, a, b = first_diff(to, te)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} emul {b}"
    assert len(to) == len(te)
    assert (ro["steps"] == re["steps"]).all()
    assert (ro["scores"] == re["scores"]).all()
    assert (ro["ranks"] == re["ranks"]).all()


def test_agari_guard_parity():
    """mortal.rs:319-336 + agent_helper.rs:262-368: with the rule-based agari guard on for every seat the decisions
    (incl. refused wins at all-last) and final scores still match the oracle."""
    n = 96
    nonces = np.arange(4000, 4000 + n, dtype=np.uint64)
    keys = np.full(n, 5, dtype=np.uint64)
    ro = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=True, trace_cap=1 << 18)
    re = E.run(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=True, trace_cap=1 << 18)
    assert (re["errs"] == 0).all(), re["errs"]
    rn = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=False)
    assert (rn["scores"] != ro["scores"]).any(), "guard never fired: the test does not cover it"
    to, te = sort_trace(ro["trace"]), sort_trace(re["trace"])
    i, a, b = first_diff(to, te)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} emul {b}"
    assert (ro["scores"] == re["scores"]).all() and (ro["ranks"] == re["ranks"]).all()


def test_emul_shanten_matches_oracle_random_hands():
    rng = np.random.default_rng(0)
    n = 20000
    tiles = np.zeros((n, 34), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint8)
    deck = np.repeat(np.arange(34, dtype=np.uint8), 4)
    for i in range(n):
        k = rng.integers(0, 4) if i % 4 == 0 else 0
        cnt = 13 + (i & 1) - 3 * k
        pick = rng.permutation(deck)[:cnt]
        np.add.at(tiles[i], pick, 1)
        lens[i] = 4 - k
    out = np.zeros(n, dtype=np.int8)
    E.lib().emul_shanten(tiles.ctypes.data, lens.ctypes.data, out.ctypes.data, n)
    assert (out == O.shanten(tiles, lens)).all()


def test_emul_wall_matches_oracle():
    for kind in (0, 1):
        for kyoku, honba in ((0, 0), (3, 2), (7, 0), (11, 5)):
            a = np.zeros(136, dtype=np.uint8)
            b = np.zeros(136, dtype=np.uint8)
            E.lib().emul_make_wall(12345678901234567, 0xDEADBEEFCAFE, kyoku, honba, kind, a.ctypes.data)
            O.lib().orc_make_wall(12345678901234567, 0xDEADBEEFCAFE, kyoku, honba, kind, b.ctypes.data)
            assert (a == b).all()


def test_obs_encode_parity_emulated():
    """v4 observation rows 0..888 + legal masks of the product encoder (host-emulated) vs the oracle."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs()
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=6, min_rows=1500)


@pytest.mark.parametrize("version", [1, 2, 3])
def test_obs_encode_parity_emulated_legacy_versions(version):
    """obs versions 1-3 (consts.rs:20-28; 938 / 942 / 934 rows) of the product encoder (host-emulated) vs the oracle."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs(version=version)
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=4, min_rows=1000, version=version)


def test_sp_block_parity_emulated():
    """obs v4 rows 889..1011 (single-player tables): product DP (host-emulated) vs the oracle's memoised recursion.
    Same f32 operation order on both sides, so the comparison is exact."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs(sp=True)
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=4, max_cycles=150, min_rows=400, sp=True, sp_tol=0.0)
    assert E.lib().emul_sp_overflows() == 0


def _log_parity(env_cls, n, policy_kind, shuffle_kind, seed0, **env_kw):
    """Drive an env and the oracle in lock step with the same decisions; compare the complete mjai event logs."""
    import json

    from mortal_b200 import mjai_log

    nonces = np.arange(seed0, seed0 + n, dtype=np.uint64)
    keys = np.full(n, 4242, dtype=np.uint64)
    env = env_cls(nonces, keys, shuffle_kind=shuffle_kind, **env_kw)
    env.enable_log()
    L = O.lib()
    games = [L.orc_game_new(int(nonces[t]), int(keys[t]), shuffle_kind, t) for t in range(n)]
    try:
        actions = None
        for cycle in range(4000):
            rt, rs, acts = env.step_and_policy(actions, policy_kind)
            for t in range(n):
                assert L.orc_game_poll(games[t]) >= 0, O.err()
            chosen = {(int(rt[r]), int(rs[r] & 3), bool(rs[r] & 4)): int(acts[r]) for r in range(len(rt))}
            for (t, seat, kan), a in chosen.items():
                if kan:
                    continue
                ka = chosen.get((t, seat, True), -1)
                assert L.orc_game_set_action(games[t], seat, a, ka if a == 42 else -1) == 0, O.err()
            for t in range(n):
                L.orc_game_advance_step(games[t])
            actions = acts
            if env.num_live() == 0:
                break
        else:
            raise AssertionError("games did not finish")
        words, lens = env.read_log()
        n_events = 0
        for t in range(n):
            ours = mjai_log.decode_events(words[t, : int(lens[t])])
            buf = (O.OrcEvent * 8192)()
            m = L.orc_game_log(games[t], buf, 8192)
            assert m <= 8192
            ref = [O.event_to_dict(buf[i]) for i in range(m)]
            assert len(ours) == len(ref), (t, len(ours), len(ref))
            for i, (a, b) in enumerate(zip(ours, ref)):
                assert a == b, (t, i, a, b)
            # and the serialised form is what serde writes: compact separators, `type` first
            text = mjai_log.dump_json_log(ours, ["a", "b", "c", "d"], (int(nonces[t]), int(keys[t])))
            lines = text.strip().split("\n")
            assert json.loads(lines[0]) == {"type": "start_game", "names": ["a", "b", "c", "d"], "seed": [int(nonces[t]), 4242]}
            assert lines[0].startswith('{"type":"start_game","names":["a","b","c","d"],"seed":[')
            assert lines[-1] == '{"type":"end_game"}' and len(lines) == len(ours) + 2
            n_events += len(ours)
        return n_events
    finally:
        for g in games:
            L.orc_game_free(g)
        env.close()


class _EmulLogEnv(E.EmulEnv):
    def step_and_policy(self, actions, kind):
        self.step(actions)
        rt, rs, _ = self.rows()
        return rt, rs, self.policy_test(kind)


@pytest.mark.parametrize("policy_kind,shuffle_kind", [(1, 0), (0, 1)])
def test_mjai_event_log_parity_emulated(policy_kind, shuffle_kind):
    """the product's device-side event log, decoded by mortal_b200.mjai_log, equals the oracle's
    game log event for event (start_kyoku with haipai, draws, calls, dora, riichi, hora with deltas and ura markers,
    ryukyoku, end_kyoku) over whole hanchans."""
    n_events = _log_parity(_EmulLogEnv, 12, policy_kind, shuffle_kind, 31337, enable_quick_eval=False)  # the oracle stepping API has no quick-eval
    assert n_events > 12 * 500


def _selfplay_logs(n, policy_kind, seed0):
    """whole-hanchan mjai logs (lists of event dicts) produced by the emulated env under a test policy"""
    from mortal_b200 import mjai_log

    nonces = np.arange(seed0, seed0 + n, dtype=np.uint64)
    keys = np.full(n, 11, dtype=np.uint64)
    env = E.EmulEnv(nonces, keys, enable_quick_eval=True)
    env.enable_log()
    acts = None
    for _ in range(4000):
        env.step(acts)
        acts = env.policy_test(policy_kind)
        if env.num_live() == 0:
            break
    words, lens = env.read_log()
    env.close()
    return [[{"type": "start_game", "names": ["a", "b", "c", "d"], "seed": [int(nonces[t]), 11]}]
            + mjai_log.decode_events(words[t, : int(lens[t])]) + [{"type": "end_game"}] for t in range(n)]


@pytest.mark.parametrize("policy_kind", [1, 0])
def test_log_replay_labels_match_gameplay_loader_emulated(policy_kind):
    """the product's log replay (csrc/mjx_replay.cuh, host-emulated) against the oracle's restatement of
    dataset/gameplay.rs: per (game, player) the same moves with the same labels, masks, at_kyoku, at_turn, shanten,
    apply_gamma, including kan-select rows."""
    from mortal_b200 import dataset_codec as DC

    games = _selfplay_logs(6, policy_kind, 900)
    jobs = DC.build_jobs(games, [[0, 1, 2, 3]] * len(games))
    rep = E.EmulReplay(jobs)
    per_job = [dict(actions=[], masks=[], meta=[], kan=0) for _ in range(rep.n_tables)]
    for _ in range(3000):
        rep.replay_step()
        if rep.num_rows():
            rt, rs, m = rep.rows()
            lab, meta = rep.row_labels()
            for r in range(len(rt)):
                j = per_job[rt[r]]
                j["actions"].append(int(lab[r])); j["masks"].append(m[r]); j["meta"].append(meta[r].copy()); j["kan"] += int(rs[r] >> 2) & 1
        if rep.live == 0:
            break
    assert rep.live == 0 and (rep.errs() == 0).all()
    moves = 0
    for job in range(rep.n_tables):
        ref = O.gameplay_load(games[jobs["job_game"][job]], int(jobs["players"][job]), with_obs=False, sp_mode=0)
        got = per_job[job]
        assert got["actions"] == ref["actions"].tolist(), job
        meta = np.array(got["meta"])
        assert (np.array(got["masks"]) == ref["masks"]).all(), job
        assert (meta[:, 0] == ref["at_kyoku"]).all() and (meta[:, 1] == ref["at_turns"]).all(), job
        assert (meta[:, 2].astype(np.int8) == ref["shantens"]).all() and (meta[:, 3].astype(bool) == ref["apply_gamma"]).all(), job
        moves += len(got["actions"])
    assert moves > 4000 and sum(j["kan"] for j in per_job) > 0
    rep.close()


def test_log_replay_observations_match_gameplay_loader_emulated():
    """the observations emitted during replay equal the oracle loader's (v4 incl. the single-player block, v3, v1)"""
    from mortal_b200 import dataset_codec as DC

    games = _selfplay_logs(2, 1, 1900)
    for version, sp in ((4, True), (3, False), (1, False)):
        jobs = DC.build_jobs(games, [[0, 2], [1, 3]])
        rep = E.EmulReplay(jobs)
        obs_per = [[] for _ in range(rep.n_tables)]
        for _ in range(3000):
            rep.replay_step()
            if rep.num_rows():
                rt, _, _ = rep.rows()
                obs = rep.encode_obs(sp=sp, version=version)
                for r in range(len(rt)):
                    obs_per[rt[r]].append(obs[r])
            if rep.live == 0:
                break
        for job in range(rep.n_tables):
            ref = O.gameplay_load(games[jobs["job_game"][job]], int(jobs["players"][job]), version=version, sp_mode=1 if sp else 0)
            got = np.array(obs_per[job])
            assert got.shape == ref["obs"].shape
            d = np.abs(got - ref["obs"])
            assert not ((d != 0) & ((ref["obs"] == 0) | (ref["obs"] == 1))).any() and d.max() <= 1e-6
        rep.close()


@pytest.mark.parametrize("quick_eval,policy_kind", [(False, 0), (True, 1)])
def test_log_meta_attachment_emulated(quick_eval, policy_kind):
    """mortal_b200.mjai_log.attach_meta: with the per-step log bounds every agent event gets exactly the decision that caused it
    (its action decodes to that event), pass decisions and overridden calls attach to nothing, and with quick-eval the events
    that had no decision row stay without meta, as in the reference (mortal.rs:161-186, 210-242)."""
    from mortal_b200 import mjai_log

    n = 10
    nonces = np.arange(2600, 2600 + n, dtype=np.uint64)
    keys = np.full(n, 8, dtype=np.uint64)
    env = E.EmulEnv(nonces, keys, enable_quick_eval=quick_eval)
    env.enable_log()
    bounds, decisions = [], [dict() for _ in range(n)]
    acts = None
    for cyc in range(4000):
        env.step(acts)
        bounds.append(env.log_lens().copy())
        rt, rs, masks = env.rows()
        acts = env.policy_test(policy_kind)
        rows = {}
        for r in range(len(rt)):
            rows[(int(rt[r]), int(rs[r] & 3), bool(rs[r] & 4))] = r
        for (t, seat, kan), r in rows.items():
            if kan:
                continue
            q = np.where(masks[r], np.float32(0.25) * np.arange(46, dtype=np.float32), -np.inf)
            kr = rows.get((t, seat, True))
            kan_meta = None
            if kr is not None and int(acts[r]) == 42:
                kan_meta = {k: v for k, v in mjai_log.make_meta(int(acts[kr]), masks[kr], q).items() if not k.startswith("_") and v is not None}
            decisions[t].setdefault(cyc, {})[seat] = mjai_log.make_meta(int(acts[r]), masks[r], q, kan_select=kan_meta)
        if env.num_live() == 0:
            break
    words, lens = env.read_log()
    env.close()
    bounds = np.array(bounds)
    tile_id = {name: i for i, name in enumerate(mjai_log.TILE_NAMES)}
    total_events = total_meta = 0
    for t in range(n):
        events, offsets = mjai_log.decode_events(words[t, : int(lens[t])], with_offsets=True)
        got = mjai_log.attach_meta(events, offsets, [int(b) for b in bounds[:, t]], decisions[t])
        agent = [e for e in events if e["type"] in mjai_log.AGENT_EVENT_ACTIONS and (e["type"] != "ryukyoku" or "meta" in e)]
        with_meta = [e for e in agent if "meta" in e]
        assert got == len(with_meta)
        n_applied = sum(1 for c in decisions[t].values() for m in c.values() if m["_action"] != 45)
        # every decision row that was not a pass produced at most one event; calls overridden by a higher-priority reaction none
        assert len(with_meta) <= n_applied
        if not quick_eval:
            unmatched = [e for e in agent if "meta" not in e]
            assert not unmatched, unmatched[:3]  # no quick-eval: every agent event has a decision behind it
        for e in with_meta:
            m = e["meta"]
            assert bin(m["mask_bits"]).count("1") == len(m["q_values"]) and m["is_greedy"] is True and "_action" not in m
            if e["type"] == "dahai":  # the legal mask of the decision allows exactly this discard
                assert (m["mask_bits"] >> tile_id[e["pai"]]) & 1
            if e["type"] in ("ankan", "kakan") and not quick_eval:  # with quick-eval a single kan candidate needs no second row
                assert "kan_select" in m and bin(m["kan_select"]["mask_bits"]).count("1") >= 1
        # the serialised line keeps serde's field order: event fields, then meta
        line = json.dumps(with_meta[0], separators=(",", ":"))
        assert line.startswith('{"type":"') and ',"meta":{"q_values":[' in line
        total_events += len(agent)
        total_meta += len(with_meta)
    assert total_meta > 3000 and (quick_eval or total_meta == total_events)
    if quick_eval:
        assert total_meta < total_events  # some discards were forced and never reached the engine


def test_log_replay_with_augmentation_emulated():
    """gameplay.rs:126-128 + mjai/event.rs:187-217: the manzu <-> pinzu augmentation is an involution on the events, and the
    replay of the augmented game matches the oracle loader on the same augmented events (labels move with the tiles)."""
    from mortal_b200 import dataset_codec as DC

    games = _selfplay_logs(3, 1, 5100)
    aug = [DC.augment_events(ev) for ev in games]
    assert [DC.augment_events(ev) for ev in aug] == games and aug != games
    sk = next(e for e in games[0] if e["type"] == "start_kyoku")
    ak = next(e for e in aug[0] if e["type"] == "start_kyoku")
    assert ak["bakaze"] == sk["bakaze"] and ak["scores"] == sk["scores"]
    assert [[t[0] + {"m": "p", "p": "m"}.get(t[1], t[1]) + t[2:] if t[0].isdigit() else t for t in hand] for hand in sk["tehais"]] == ak["tehais"]
    jobs = DC.build_jobs(aug, [[0, 1, 2, 3]] * len(aug))
    rep = E.EmulReplay(jobs)
    per_job = [[] for _ in range(rep.n_tables)]
    for _ in range(3000):
        rep.replay_step()
        if rep.num_rows():
            rt, _, _ = rep.rows()
            lab, _ = rep.row_labels()
            for r in range(len(rt)):
                per_job[rt[r]].append(int(lab[r]))
        if rep.live == 0:
            break
    assert (rep.errs() == 0).all()
    swap = lambda a: a + 9 if a < 9 else a - 9 if a < 18 else {34: 35, 35: 34}.get(a, a)
    for job in range(rep.n_tables):
        g, pid = int(jobs["job_game"][job]), int(jobs["players"][job])
        ref_aug = O.gameplay_load(aug[g], pid, with_obs=False, sp_mode=0)["actions"].tolist()
        ref_raw = O.gameplay_load(games[g], pid, with_obs=False, sp_mode=0)["actions"].tolist()
        assert per_job[job] == ref_aug
        # discards (and kan-select tiles) move with the suits, everything else keeps its label
        assert len(ref_aug) == len(ref_raw) and all(b == a or b == swap(a) for a, b in zip(ref_raw, ref_aug))
    rep.close()


def test_bench_engine_plays_policy_kind_2():
    """bench.py's MaskHashEngine (a reference-protocol react_batch engine in numpy) chooses, from the legal mask and the v4
    observation planes alone, exactly what the test policy kind 2 chooses on the device / in the oracle: the CPU arm, env_only
    and e2e of the benchmark therefore play the same games."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    eng = bench.MaskHashEngine()
    n = 24
    nonces = np.arange(3000, 3000 + n, dtype=np.uint64)
    keys = np.full(n, 11, dtype=np.uint64)
    env = E.EmulEnv(nonces, keys, enable_quick_eval=True)
    env.step(None)
    checked = 0
    for cycle in range(220):
        rt, rs, masks = env.rows()
        want = env.policy_test(2)
        if len(rt):
            obs = env.encode_obs(sp=False)
            got, q, m, greedy = eng.react_batch([obs[i] for i in range(len(rt))], [masks[i] for i in range(len(rt))], None)
            assert list(got) == want[: len(rt)].tolist(), (cycle, np.nonzero(np.array(got) != want[: len(rt)])[0][:5])
            assert len(q) == len(rt) and len(q[0]) == 46 and list(m[0]) == masks[0].tolist() and all(greedy)
            checked += len(rt)
        env.step(want)
        if env.num_live() == 0:
            break
    env.close()
    assert checked > 3000


@pytest.mark.parametrize("version", [4, 1])
def test_invisible_obs_parity_emulated(version):
    """arena/board.rs:680-782 encode_oracle_obs: the device encoder (host-emulated) against the oracle's restatement at every
    decision of seeded games played in lock step (hands / akas / shanten / waits / furiten of the three other seats, the
    remaining wall in drawing order, rinshan, all dora and ura indicators)."""
    n = 6
    nonces = np.arange(880, 880 + n, dtype=np.uint64)
    keys = np.full(n, 5, dtype=np.uint64)
    env = E.EmulEnv(nonces, keys, enable_quick_eval=False)
    L = O.lib()
    games = [L.orc_game_new(int(nonces[t]), int(keys[t]), 0, t) for t in range(n)]
    rows = L.orc_oracle_obs_rows(version)
    checked = nonzero_wall = 0
    try:
        env.step(None)
        for cycle in range(260):
            rt, rs, masks = env.rows()
            inv = env.encode_invisible(version)
            acts = env.policy_test(1)
            for t in range(n):
                assert L.orc_game_poll(games[t]) >= 0, O.err()
            chosen = {}
            for r in range(len(rt)):
                t, seat, kan = int(rt[r]), int(rs[r] & 3), bool(rs[r] & 4)
                ref = np.zeros((rows, 34), dtype=np.float32)
                assert L.orc_game_encode_oracle_obs(games[t], seat, version, ref.ctypes.data) == 0, O.err()
                assert inv[r].shape == ref.shape
                bad = np.argwhere(inv[r] != ref)
                assert len(bad) == 0, (version, cycle, t, seat, bad[:6], inv[r][tuple(bad[0])], ref[tuple(bad[0])])
                checked += 1
                nonzero_wall += int(ref[51 if version != 1 else 45:].sum() > 0)
                chosen[(t, seat, kan)] = int(acts[r])
            for (t, seat, kan), a in chosen.items():
                if kan:
                    continue
                ka = chosen.get((t, seat, True), -1)
                assert L.orc_game_set_action(games[t], seat, a, ka if a == 42 else -1) == 0, O.err()
            for t in range(n):
                L.orc_game_advance_step(games[t])
            env.step(acts)
            if env.num_live() == 0:
                break
        assert checked > 1200 and nonzero_wall == checked
    finally:
        for g in games:
            L.orc_game_free(g)
        env.close()


@pytest.mark.parametrize("trust_seed", [True, False])
def test_log_replay_invisible_obs_match_gameplay_loader_emulated(trust_seed):
    """GameplayLoader(oracle=True): the invisible observation of every move (dataset/invisible.rs Invisible::encode: every tile
    left in the live wall, not just `tiles_left` of them) — with `trust_seed` the walls are regenerated from the game seed on
    both sides; without it they are reconstructed from the log with a random fill of the unseen tiles (mortal_b200.dataset_codec
    restates Invisible::new) and handed to both sides."""
    from mortal_b200 import dataset_codec as DC

    games = _selfplay_logs(4, 0, 1500)  # uniform policy: kans (rinshan draws) occur
    rng = np.random.default_rng(3)
    walls = None if trust_seed else [DC.reconstruct_walls(ev, rng) for ev in games]
    if walls is not None:  # the reconstruction keeps what the log shows and fills the rest with exactly the unseen tiles
        for ev, w in zip(games, walls):
            assert len(w) == sum(e["type"] == "start_kyoku" for e in ev)
            for q in range(len(w)):
                assert np.bincount(w[q], minlength=37).tolist() == DC.new_unknown_tiles()
    jobs = DC.build_jobs(games, [[0, 1, 2, 3]] * len(games), walls)
    rep = E.EmulReplay(jobs)
    if trust_seed:
        seeds = [ev[0]["seed"] for ev in games]
        rep.trust_seeds([seeds[g][0] for g in jobs["job_game"]], [seeds[g][1] for g in jobs["job_game"]])
    inv_per = [[] for _ in range(rep.n_tables)]
    for _ in range(3000):
        rep.replay_step()
        if rep.num_rows():
            rt, _, _ = rep.rows()
            inv = rep.encode_invisible(4)
            for r in range(len(rt)):
                inv_per[rt[r]].append(inv[r])
        if rep.live == 0:
            break
    assert rep.live == 0 and (rep.errs() == 0).all()
    kans = 0
    for job in range(rep.n_tables):
        g = jobs["job_game"][job]
        ref = O.gameplay_load(games[g], int(jobs["players"][job]), with_obs=False, sp_mode=0,
                              oracle_seed=tuple(games[g][0]["seed"]) if trust_seed else None, walls=None if trust_seed else walls[g])
        got = np.array(inv_per[job])
        assert got.shape == ref["invisible"].shape and got.shape[0] > 50
        bad = np.argwhere(got != ref["invisible"])
        assert len(bad) == 0, (job, bad[:5])
        kans += int((ref["actions"] == 42).sum())
    assert kans > 0
    rep.close()
    # a log replayed with the wrong seed fails loudly
    if trust_seed:
        rep = E.EmulReplay(DC.build_jobs(games[:1], [[0]]))
        rep.trust_seeds([12345], [6789])
        rep.replay_step()
        assert (rep.errs() != 0).any()
        rep.close()


def test_grp_features_from_table_state_equal_log_arithmetic_emulated():
    """the per-kyoku GRP feature rows (dataset/grp.rs:134-147) written by the step code when a kyoku starts
    equal what dataset.Grp computes from the game's mjai log, and the final ranking / scores it derives equal the env's results."""
    from mortal_b200 import mjai_log
    from mortal_b200.dataset import Grp

    n = 8
    nonces = np.arange(300, 300 + n, dtype=np.uint64)
    keys = np.full(n, 9, dtype=np.uint64)
    env = E.EmulEnv(nonces, keys)
    env.enable_log()
    env.enable_grp()
    acts = None
    for _ in range(4000):
        env.step(acts)
        acts = env.policy_test(1)
        if env.num_live() == 0:
            break
    words, lens = env.read_log()
    feats = env.read_grp()
    res = E.results(env) if hasattr(E, "results") else None
    for t in range(n):
        ev = [{"type": "start_game", "names": list("abcd")}] + mjai_log.decode_events(words[t, : int(lens[t])]) + [{"type": "end_game"}]
        g = Grp.load_events(ev)
        assert g.feature.shape == feats[t].shape and (g.feature == feats[t]).all() and feats[t].dtype == np.float64
    env.close()
