"""GPU parity tests: the CUDA path through the C ABI (mortal_b200/libmjx.so) against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from test_emul_vs_oracle import first_diff, sort_trace

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mjx():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import mortal_b200
    from mortal_b200 import _lib

    _lib.init(0)
    return mortal_b200


def random_hands(n, seed=0):
    rng = np.random.default_rng(seed)
    tiles = np.zeros((n, 34), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint8)
    deck = np.repeat(np.arange(34, dtype=np.uint8), 4)
    for i in range(n):
        k = int(rng.integers(0, 4)) if i % 4 == 0 else 0
        cnt = 13 + (i & 1) - 3 * k
        pick = rng.permutation(deck)[:cnt]
        np.add.at(tiles[i], pick, 1)
        lens[i] = 4 - k
    return tiles, lens


def test_shanten_kernel_bit_exact(mjx):
    from mortal_b200 import _lib

    L = _lib.load()
    tiles, lens = random_hands(100_003)
    out = np.zeros(len(lens), dtype=np.int8)
    _lib.check(L.mjx_shanten_host(tiles.ctypes.data, lens.ctypes.data, out.ctypes.data, len(lens)), "mjx_shanten_host")
    ref = O.shanten(tiles, lens)
    assert (out == ref).all()
    # empty and tiny inputs
    _lib.check(L.mjx_shanten_host(tiles.ctypes.data, lens.ctypes.data, out.ctypes.data, 0), "empty")
    _lib.check(L.mjx_shanten_host(tiles.ctypes.data, lens.ctypes.data, out.ctypes.data, 1), "one")
    assert out[0] == ref[0]


def winning_hands(n, seed=1):
    """Constructed winning hands (4 mentsu + pair / chiitoi / kokushi) with random melds, winds, ron/tsumo."""
    rng = np.random.default_rng(seed)
    q = np.zeros(n, dtype=O.AGARI_IN_DTYPE)
    for i in range(n):
        while True:
            counts = np.zeros(34, dtype=np.int64)
            kind = rng.integers(0, 20)
            chis, pons, minkans, ankans = [], [], [], []
            if kind == 0:  # chiitoi
                for t in rng.choice(34, 7, replace=False):
                    counts[t] += 2
            elif kind == 1:  # kokushi
                yao = [0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33]
                for t in yao:
                    counts[t] += 1
                counts[rng.choice(yao)] += 1
            else:
                closed = counts.copy()
                total = counts.copy()
                ok = True
                for _ in range(4):
                    melded = rng.random() < 0.3
                    if rng.random() < 0.55:
                        s = int(rng.integers(0, 3)) * 9 + int(rng.integers(0, 7))
                        total[s:s + 3] += 1
                        if melded:
                            chis.append(s)
                        else:
                            closed[s:s + 3] += 1
                    else:
                        t = int(rng.integers(0, 34))
                        r = rng.random()
                        if melded and r < 0.15:
                            total[t] += 4; minkans.append(t)
                        elif melded and r < 0.3:
                            total[t] += 4; ankans.append(t)
                        elif melded:
                            total[t] += 3; pons.append(t)
                        else:
                            total[t] += 3; closed[t] += 3
                p = int(rng.integers(0, 34))
                total[p] += 2
                closed[p] += 2
                if (total > 4).any():
                    continue
                counts = closed
            if (counts > 4).any():
                continue
            break
        q["tehai"][i] = counts
        for name, v in (("chis", chis), ("pons", pons), ("minkans", minkans), ("ankans", ankans)):
            q[name][i][: len(v)] = v
            q["n_" + name][i] = len(v)
        present = np.nonzero(counts)[0]
        q["winning_tile"][i] = rng.choice(present)
        q["bakaze"][i] = 27 + rng.integers(0, 3)
        q["jikaze"][i] = 27 + rng.integers(0, 4)
        q["is_ron"][i] = rng.integers(0, 2)
        q["additional_hans"][i] = rng.integers(0, 4)
        q["doras"][i] = rng.integers(0, 5)
        q["is_oya"][i] = q["jikaze"][i] == 27
    return q


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_agari_kernel_bit_exact(mjx, mode):
    from mortal_b200 import _lib

    L = _lib.load()
    q = winning_hands(30_000)
    out = np.zeros(len(q), dtype=O.AGARI_OUT_DTYPE)
    _lib.check(L.mjx_agari_host(q.ctypes.data, out.ctypes.data, len(q), mode), "mjx_agari_host")
    ref = O.agari(q, mode)
    for f in ("kind", "fu", "han", "yakuman", "ron", "tsumo_ko", "tsumo_oya"):
        bad = np.nonzero(out[f] != ref[f])[0]
        assert len(bad) == 0, (f, bad[:5], out[bad[:5]], ref[bad[:5]])
    assert (ref["kind"] != 0).mean() > 0.4


def test_wall_kernel_matches_oracle(mjx):
    from mortal_b200 import _lib

    L = _lib.load()
    for kind in (0, 1):
        for kyoku, honba in ((0, 0), (5, 3), (11, 9)):
            a = np.zeros(136, dtype=np.uint8)
            b = np.zeros(136, dtype=np.uint8)
            _lib.check(L.mjx_make_wall_host(10637, 12210010324280706444, kyoku, honba, kind, a.ctypes.data), "wall")
            O.lib().orc_make_wall(10637, 12210010324280706444, kyoku, honba, kind, b.ctypes.data)
            assert (a == b).all()


@pytest.mark.parametrize("policy_kind,quick_eval,shuffle_kind", [(1, True, 0), (0, False, 1), (1, False, 0), (0, True, 1)])
def test_selfplay_trace_parity(mjx, policy_kind, quick_eval, shuffle_kind):
    """BASELINE config 1 shape: OneVsThree seed layout, counter-based agents; every decision row and the
    end-of-hanchan scores / rankings bit-equal to the oracle."""
    n = 64
    nonces = np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4)
    keys = np.full(n, 0x2000, dtype=np.uint64)
    env = mjx.BatchEnv(nonces, keys, shuffle_kind=shuffle_kind, enable_quick_eval=quick_eval)
    got = env.run_test_policy(kind=policy_kind, trace=True)
    ref = O.run_batch(nonces, keys, shuffle_kind=shuffle_kind, policy_kind=policy_kind, quick_eval=quick_eval,
                      trace_cap=1 << 18)
    assert (got["err"] == 0).all(), got["err"]
    assert (got["done"] == 1).all()
    to, tg = sort_trace(ref["trace"]), sort_trace(got["trace"])
    i, a, b = first_diff(to, tg)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} gpu {b}"
    assert len(to) == len(tg)
    assert (got["steps"] == ref["steps"]).all()
    assert (got["scores"] == ref["scores"]).all()
    assert (got["ranks"] == ref["ranks"]).all()
    env.close()


def test_selfplay_4096_tables_scores(mjx):
    """BASELINE config 2 size (4096 tables): scores/ranks/steps equal to the oracle; sum of scores conserved."""
    n = 4096
    nonces = np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4)
    keys = np.full(n, 0x2000, dtype=np.uint64)
    env = mjx.BatchEnv(nonces, keys)
    got = env.run_test_policy(kind=1)
    assert (got["err"] == 0).all() and (got["done"] == 1).all()
    assert (got["scores"].sum(axis=1) == 100000).all()
    ref = O.run_batch(nonces, keys, policy_kind=1, n_threads=8)
    assert (got["scores"] == ref["scores"]).all()
    assert (got["ranks"] == ref["ranks"]).all()
    assert (got["steps"] == ref["steps"]).all()
    assert env.total_steps() == int(ref["steps"].sum())
    env.close()


def test_obs_encode_matches_oracle(mjx):
    """All 1012 rows of the v4 observation (incl. the single-player block 889-1011) and the legal mask, compared
    at every decision of a few seeded games. Everything exactly, except the exp() planes (<= 1e-6)."""
    import torch

    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        env = mjx.BatchEnv(nonces, keys, enable_quick_eval=False)
        env._actions = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
        orig_close = env.close

        def close_and_record():
            if env._h:
                make_env.overflows = env.sp_overflows()
            orig_close()

        env.close = close_and_record
        return env

    def fetch(env, first, prev):
        env.step(None if first else env._actions)
        obs = env.encode_obs()
        env.policy_test(1, env._actions)
        nr = env.num_rows()
        return (env.row_table[:nr].cpu().numpy(), env.row_seat[:nr].cpu().numpy(), env.masks[:nr].cpu().numpy(),
                obs[:nr].cpu().numpy(), env._actions[:nr].cpu().numpy())

    check_obs_parity(make_env, fetch, sp=True, sp_tol=0.0)
    assert make_env.overflows == 0


@pytest.mark.parametrize("version", [1, 2, 3])
def test_obs_encode_legacy_versions_match_oracle(mjx, version):
    """obs versions 1-3 on device vs the oracle (exact on 0/1 cells, 1e-6 on the exp()-derived planes)."""
    import torch

    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        env = mjx.BatchEnv(nonces, keys, enable_quick_eval=False, obs_version=version)
        env._actions = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
        return env

    def fetch(env, first, prev):
        env.step(None if first else env._actions)
        obs = env.encode_obs()
        env.policy_test(1, env._actions)
        nr = env.num_rows()
        return (env.row_table[:nr].cpu().numpy(), env.row_seat[:nr].cpu().numpy(), env.masks[:nr].cpu().numpy(),
                obs[:nr].cpu().numpy(), env._actions[:nr].cpu().numpy())

    check_obs_parity(make_env, fetch, n=6, min_rows=1500, version=version)


def test_encode_obs_host_equals_device_encode(mjx):
    """mjx_env_encode_obs_host (D2H overlapped with the SP kernels) delivers exactly the device encoding."""
    import torch

    n = 64
    nonces = np.arange(500, 500 + n, dtype=np.uint64)
    keys = np.full(n, 3, dtype=np.uint64)
    env = mjx.BatchEnv(nonces, keys)
    actions = torch.zeros(env.row_cap, dtype=torch.int64, device=env.device)
    h_obs = torch.full((env.row_cap, 1012, 34), -1.0, dtype=torch.float32).pin_memory()
    h_masks = torch.zeros((env.row_cap, 46), dtype=torch.bool).pin_memory()
    env.step(None)
    checked = 0
    for cycle in range(60):
        nr = env.encode_obs_host(h_obs, h_masks)
        assert nr == env.num_rows()
        dev = env.encode_obs()[:nr].cpu()
        assert torch.equal(h_obs[:nr], dev)
        assert torch.equal(h_masks[:nr], env.masks[:nr].cpu())
        checked += nr
        env.policy_test(1, actions)
        env.step(actions)
    assert checked > 60 * n * 0.9 and env.sp_overflows() == 0
    env.close()


def test_sp_lanes_give_the_single_dp_result(mjx, monkeypatch):
    """mjx_env_encode_obs solves the single-player block as several concurrent DPs (row shares on side streams) for large batches;
    forced on here for a small batch (MJX_SP_LANES): every observation equals the one-DP result bit for bit, no overflow."""
    import torch

    n = 96
    nonces = np.arange(900, 900 + n, dtype=np.uint64)
    keys = np.full(n, 11, dtype=np.uint64)
    envs = []
    for lanes in ("1", "3", "4"):
        monkeypatch.setenv("MJX_SP_LANES", lanes)
        envs.append(mjx.BatchEnv(nonces, keys))
    actions = [torch.zeros(env.row_cap, dtype=torch.int64, device=env.device) for env in envs]
    for env in envs:
        env.step(None)

    def ordered(env, obs):  # rows are appended by atomics: compare in (table, seat | kan-select) order
        nr = env.num_rows()
        key = env.row_table[:nr].long() * 8 + env.row_seat[:nr].long()
        return obs[:nr][torch.argsort(key)]

    checked = 0
    for cycle in range(80):
        ref = ordered(envs[0], envs[0].encode_obs())
        for env in envs[1:]:
            got = ordered(env, env.encode_obs())
            assert got.shape == ref.shape and torch.equal(got, ref)
        checked += ref.shape[0]
        for env, a in zip(envs, actions):
            env.policy_test(2, a)
            env.step(a)
    assert checked > 80 * n * 0.9
    assert all(env.sp_overflows() == 0 for env in envs)
    assert envs[2].sp_stats() == envs[0].sp_stats()  # states / edges / per-level counts summed over the lanes
    for env in envs:
        env.close()


def test_one_vs_three_network_policy_action_replay(mjx):
    """BASELINE config 2 protocol (SURVEY.md §8d ii): a float policy (random-init Mortal brain, greedy) drives the CUDA
    arena through the libriichi-compatible OneVsThree.py_vs_py; the recorded decisions are replayed in the oracle,
    which must accept every action as legal and reproduce scores, rankings and the returned histogram bit for bit."""
    import torch

    import mortal_b200.libriichi as lr
    from mortal_b200.engine import DeviceEngine
    from mortal_b200.model import DQN, Brain

    lr.install()
    from libriichi.arena import OneVsThree

    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    mk = lambda name: DeviceEngine(Brain(conv_channels=32, num_blocks=2, version=4), DQN(version=4), device=dev,
                                   enable_amp=True, enable_quick_eval=True, name=name)
    challenger, champion = mk("a"), mk("b")
    arena = OneVsThree(disable_progress_bar=True, log_dir=None)
    arena.record_decisions = True
    seed_start, seed_count = (10000, 0x2000), 6
    rankings = arena.py_vs_py(challenger=challenger, champion=champion, seed_start=seed_start, seed_count=seed_count)
    assert sum(rankings) == 4 * seed_count
    n = 4 * seed_count
    nonces = np.repeat(np.arange(seed_start[0], seed_start[0] + seed_count, dtype=np.uint64), 4)
    keys = np.full(n, seed_start[1], dtype=np.uint64)
    ref = O.run_replay(nonces, keys, arena.last_decisions, quick_eval=True)
    got = arena.last_results
    assert (got["scores"] == ref["scores"]).all()
    assert (got["ranks"] == ref["ranks"]).all()
    assert (got["steps"] == ref["steps"]).all()
    hist = [0, 0, 0, 0]
    for i in range(n):
        hist[int(ref["ranks"][i, i % 4])] += 1
    assert hist == rankings


def test_mjai_event_log_matches_oracle(mjx):
    """SURVEY.md §8f N1 on device: the event log recorded by k_step, decoded by mortal_b200.mjai_log, equals the oracle's
    game log event for event over whole hanchans (same decisions on both sides)."""
    import torch

    from test_emul_vs_oracle import _log_parity

    class GpuLogEnv(mjx.BatchEnv):
        def __init__(self, nonces, keys, **kw):
            super().__init__(nonces, keys, **kw)
            self._a = torch.zeros(self.row_cap, dtype=torch.int64, device=self.device)

        def step_and_policy(self, actions, kind):
            self.step(None if actions is None else self._a)
            self.policy_test(kind, self._a)
            nr = self.num_rows()
            return self.row_table[:nr].cpu().numpy(), self.row_seat[:nr].cpu().numpy(), self._a.cpu().numpy()

    n_events = _log_parity(GpuLogEnv, 48, 1, 0, 777, enable_quick_eval=False)
    assert n_events > 48 * 500


def test_arena_writes_mjai_logs(mjx, tmp_path):
    """OneVsThree(log_dir=...) writes one {seed}_{key}_{a|b|c|d}.json.gz per game (one_vs_three.rs:195-225) whose final
    scores are consistent with the returned results."""
    import gzip
    import json

    import torch

    from mortal_b200.libriichi.arena import OneVsThree

    class Greedy:
        engine_type = "mortal"
        version = 4
        is_oracle = False
        enable_quick_eval = True
        enable_rule_based_agari_guard = False

        def __init__(self, name):
            self.name = name

        def react_device(self, obs, masks):
            q = torch.rand(masks.shape, device=masks.device) + 10.0 * obs[:, 876, :].new_zeros(masks.shape)
            q = q.masked_fill(~masks, -1.0)
            return q.argmax(-1), q

    arena = OneVsThree(disable_progress_bar=True, log_dir=str(tmp_path))
    arena.record_decisions = True
    arena.record_grp = True
    torch.manual_seed(0)
    rankings = arena.py_vs_py(Greedy("chal"), Greedy("champ"), (4000, 77), 3)
    assert sum(rankings) == 12
    # GRP features straight from the table records == dataset.Grp over the written logs (dataset/grp.rs:90-164)
    from mortal_b200.dataset import Grp
    for g, path in enumerate(arena.last_log_paths):
        ref_g = Grp.load_gz_log_files([path])[0]
        got_g = arena.last_grp[g]
        assert (got_g.take_feature() == ref_g.take_feature()).all() and got_g.take_rank_by_player() == ref_g.take_rank_by_player()
        assert got_g.take_final_scores() == ref_g.take_final_scores()
    # agent/mortal.rs:161-186 gen_meta on the GPU path: the recorder must not have failed, and every logged agent event's meta is
    # the decision that caused it: legal-only q-values, the legal mask, shanten / furiten, is_greedy
    assert arena.last_meta_error is None, repr(arena.last_meta_error)
    dec, bits = arena.last_decisions, arena.last_decision_masks
    nonces_m = np.repeat(np.arange(4000, 4003, dtype=np.uint64), 4)
    O.run_replay(nonces_m, np.full(12, 77, dtype=np.uint64), dec, quick_eval=True, mask_bits=bits)  # the recorded masks ARE the oracle's
    mask_of = {}
    for (t, _, seat, kan, a), b in zip(dec.tolist(), bits.tolist()):
        if not kan:
            mask_of.setdefault((t, seat), []).append((a, b))
    n_meta = 0
    for g, path in enumerate(arena.last_log_paths):
        pos = {s: 0 for s in range(4)}
        for ev in (json.loads(ln) for ln in gzip.open(path, "rt")):
            m = ev.get("meta")
            if m is None:
                continue
            n_meta += 1
            assert len(m["q_values"]) == bin(m["mask_bits"]).count("1") and m["is_greedy"] is True
            assert 0 <= m["shanten"] <= 6 and isinstance(m["at_furiten"], bool) and m["batch_size"] > 0 and m["eval_time_ns"] > 0
            seat = ev.get("actor")
            if seat is None:
                continue
            # the seat's recorded decisions are in order; this event's meta must be one of them, further down the list
            lst = mask_of[(g, seat)]
            while pos[seat] < len(lst) and lst[pos[seat]][1] != m["mask_bits"]:
                pos[seat] += 1
            assert pos[seat] < len(lst), (g, seat, ev)
    assert n_meta > 12 * 100
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == sorted(f"{4000 + s}_77_{c}.json.gz" for s in range(3) for c in "abcd")
    res = arena.last_results
    for g, path in enumerate(arena.last_log_paths):
        lines = [json.loads(ln) for ln in gzip.open(path, "rt")]
        assert lines[0]["type"] == "start_game" and lines[0]["seed"] == [4000 + g // 4, 77]
        assert lines[0]["names"] == ["chal" if s == g % 4 else "champ" for s in range(4)]
        assert lines[-1] == {"type": "end_game"}
        scores, sticks = None, 0
        for ev in lines:
            if ev["type"] == "start_kyoku":
                scores, sticks = list(ev["scores"]), ev["kyotaku"]
            elif ev["type"] in ("hora", "ryukyoku"):
                scores = [a + b for a, b in zip(scores, ev["deltas"])]
                if ev["type"] == "hora":
                    sticks = 0  # the first winner's deltas already contain the sticks on the table (board.rs:401-403)
            elif ev["type"] == "reach_accepted":
                scores[ev["actor"]] -= 1000
                sticks += 1
        if sticks:  # game.rs:181-198: sticks left at the very end go to the first top seat
            scores[max(range(4), key=lambda i: (scores[i], -i))] += 1000 * sticks
        assert scores == [int(x) for x in res["scores"][g]], (g, scores, res["scores"][g])


def test_agari_guard_on_device_matches_oracle(mjx):
    """mortal.rs:319-336 + agent_helper.rs:262-368 on device: with the rule-based agari guard on for every seat the
    decisions (incl. refused wins at all-last) and final scores match the oracle, and differ from the unguarded run."""
    n = 256
    nonces = np.arange(4000, 4000 + n, dtype=np.uint64)
    keys = np.full(n, 5, dtype=np.uint64)
    ro = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=True, trace_cap=1 << 19, n_threads=8)
    rn = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=False, n_threads=8)
    assert (rn["scores"] != ro["scores"]).any(), "guard never fired: the test does not cover it"
    env = mjx.BatchEnv(nonces, keys, enable_quick_eval=True)
    rg = env.run_test_policy(1, trace=True, agari_guard=True)
    env.close()
    assert (rg["err"] == 0).all()
    to, tg = sort_trace(ro["trace"]), sort_trace(rg["trace"])
    i, a, b = first_diff(to, tg)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} gpu {b}"
    assert (ro["scores"] == rg["scores"]).all() and (ro["ranks"] == rg["ranks"]).all()


def test_two_vs_two_arena_runs_and_logs(mjx, tmp_path):
    """TwoVsTwo.py_vs_py (two_vs_two.rs:35-55) returns None; its product is the log files: 2 games per seed named
    {seed}_{key}_{a|b}.json.gz, challenger pair on seats {0,2} in game a and {1,3} in game b."""
    import gzip
    import json

    import torch

    from mortal_b200.libriichi.arena import TwoVsTwo

    class Rand:
        engine_type = "mortal"
        version = 4
        is_oracle = False
        enable_quick_eval = True
        enable_rule_based_agari_guard = False

        def __init__(self, name):
            self.name = name

        def react_device(self, obs, masks):
            q = torch.rand(masks.shape, device=masks.device).masked_fill(~masks, -1.0)
            return q.argmax(-1), q

    torch.manual_seed(1)
    arena = TwoVsTwo(disable_progress_bar=True, log_dir=str(tmp_path))
    assert arena.py_vs_py(Rand("x"), Rand("y"), (9000, 3), 8) is None
    res = arena.last_results
    assert res["ranks"].shape[0] == 16
    assert (np.sort(res["ranks"], axis=1) == np.arange(4)).all() and (res["scores"].sum(1) % 100 == 0).all()
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == sorted(f"{9000 + s}_3_{c}.json.gz" for s in range(8) for c in "ab")
    for g, path in enumerate(arena.last_log_paths):
        first = json.loads(gzip.open(path, "rt").readline())
        chal = [0, 2] if g % 2 == 0 else [1, 3]
        assert first["names"] == ["x" if s in chal else "y" for s in range(4)]


def test_reference_protocol_engine_through_host_buffers(mjx):
    """An engine that only implements the reference protocol (react_batch over lists of numpy arrays, mortal.rs:126-152)
    runs through OneVsThree unchanged; the arena feeds it from pinned host buffers (mjx_env_encode_obs_host). The recorded
    decisions replay in the oracle to the same scores, and every observation it saw had the right shape and range."""
    from mortal_b200.libriichi.arena import OneVsThree

    seen = dict(rows=0, batches=0)

    class RefProtocolEngine:
        engine_type = "mortal"
        name = "ref"
        version = 4
        is_oracle = False
        enable_quick_eval = True
        enable_rule_based_agari_guard = False

        def __init__(self, seed):
            self.rng = np.random.default_rng(seed)

        def react_batch(self, obs, masks, invisible_obs):
            assert invisible_obs is None and len(obs) == len(masks) > 0
            assert obs[0].shape == (1012, 34) and obs[0].dtype == np.float32 and masks[0].shape == (46,)
            o = np.stack(obs)
            assert o.min() >= 0.0 and o.max() <= 1.0
            m = np.stack(masks).astype(bool)
            q = self.rng.random(m.shape, dtype=np.float32)
            q[~m] = -np.inf
            seen["rows"] += len(obs)
            seen["batches"] += 1
            return q.argmax(-1).tolist(), q.tolist(), m.tolist(), [True] * len(obs)

    arena = OneVsThree(disable_progress_bar=True)
    arena.record_decisions = True
    rankings = arena.py_vs_py(RefProtocolEngine(1), RefProtocolEngine(2), (6000, 9), 4)
    assert sum(rankings) == 16 and seen["rows"] > 16 * 300
    dec = arena.last_decisions
    order = np.lexsort((dec[:, 3], dec[:, 2], dec[:, 1], dec[:, 0]))
    nonces = np.repeat(np.arange(6000, 6004, dtype=np.uint64), 4)
    keys = np.full(16, 9, dtype=np.uint64)
    ref = O.run_replay(nonces, keys, dec[order], quick_eval=True)
    assert (ref["scores"] == arena.last_results["scores"]).all() and (ref["ranks"] == arena.last_results["ranks"]).all()


def test_fused_policy_net_kernels_match_torch(mjx):
    """csrc/mjx_nn.cuh against plain PyTorch fp32 references of the same ops (bf16 in / out, fp32 math: <= 1 bf16 ulp),
    and the fused bf16 fast path of the 192-channel brain against its fp32 forward."""
    import torch

    from mortal_b200 import nn_ops
    from mortal_b200.model import Brain

    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    B, Cc, L = 257, 192, 34
    x = (torch.randn(B, Cc, 1, L, device=dev) * 2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = (torch.randn(B, Cc, 1, L, device=dev) * 2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    scale = torch.rand(Cc, device=dev) + 0.5
    bias = torch.randn(Cc, device=dev)
    ref = torch.nn.functional.mish(x.float() * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    got = nn_ops.affine_mish(x, scale, bias).float()
    assert (got - ref).abs().max() <= 2 ** -7 * ref.abs().max()  # one bf16 rounding (8 significant bits)
    assert ((got - ref).abs() <= ref.abs() * 2 ** -8 * 1.01 + 1e-5).all()  # (SFU exp / reciprocal: ~1e-6 before the rounding)
    avg, mx = nn_ops.pool_mean_max(x)
    assert ((avg.float() - x.float().mean((2, 3))).abs() <= x.float().mean((2, 3)).abs() * 2 ** -8 + 1e-3).all()
    assert torch.equal(mx.float(), x.float().amax((2, 3)))
    gate = torch.rand(B, Cc, device=dev).to(torch.bfloat16)
    ref = y.float() * gate.float().view(B, Cc, 1, 1) + x.float()
    got = nn_ops.gate_residual(y, gate, x).float()
    assert ((got - ref).abs() <= ref.abs() * 2 ** -8 + 1e-6).all()

    # the fused block tail (pool -> gate MLP -> sigmoid -> y * gate + x -> next BN-affine + Mish) against an fp32 composition
    H = Cc // 16
    w1 = (torch.randn(H, Cc, device=dev) * 0.3).to(torch.bfloat16).float()
    b1 = (torch.randn(H, device=dev) * 0.3).to(torch.bfloat16).float()
    w2 = (torch.randn(Cc, H, device=dev) * 0.8).to(torch.bfloat16).float()
    b2 = (torch.randn(Cc, device=dev) * 0.3).to(torch.bfloat16).float()
    mlp = lambda v: torch.nn.functional.mish(v @ w1.T + b1) @ w2.T + b2
    yf, xf = y.float(), x.float()
    gate32 = torch.sigmoid(mlp(yf.mean((2, 3))) + mlp(yf.amax((2, 3))))
    assert gate32.min() < 0.15 and gate32.max() > 0.85  # a gate that varies: an indexing slip cannot hide in the tolerance
    ref_x = yf * gate32.view(B, Cc, 1, 1) + xf
    got_x, got_a = nn_ops.block_tail(y, x, w1, b1, w2.T.contiguous(), b2, scale, bias)
    assert got_x.shape == y.shape and got_x.is_contiguous(memory_format=torch.channels_last)
    # the gate is computed in fp32 and stored as bf16: within one bf16 rounding (2^-8 relative, gate <= 1) plus accumulation noise
    gate_err = ((got_x.float() - ref_x).abs() - ref_x.abs() * 2 ** -7 - 1e-3) / yf.abs().clamp_min(1e-3)
    assert gate_err.max().item() <= 0.006, gate_err.max().item()
    ref_a = torch.nn.functional.mish(got_x.float() * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    assert ((got_a.float() - ref_a).abs() <= ref_a.abs() * 2 ** -8 * 1.01 + 1e-5).all()
    # and against the unfused kernels it replaces: the same up to the rounding of the gate
    avg_y, mx_y = nn_ops.pool_mean_max(y)
    hb = lambda v: (torch.nn.functional.mish((v @ w1.to(torch.bfloat16).T + b1.to(torch.bfloat16))) @ w2.to(torch.bfloat16).T + b2.to(torch.bfloat16))
    gate_b = torch.sigmoid(hb(avg_y) + hb(mx_y)).contiguous()
    old_x = nn_ops.gate_residual(y, gate_b, x)
    dev_old = ((got_x.float() - old_x.float()).abs() - old_x.float().abs() * 2 ** -7 - 1e-3) / yf.abs().clamp_min(1e-3)
    assert dev_old.max().item() <= 0.2, dev_old.max().item()  # that pipeline rounds pooled vectors, hidden layer and logits to bf16
    assert dev_old.median().item() <= 0.0

    # the stem's input transform: f32 [B, C, L] -> bf16 channels-last, channels zero-padded to a multiple of 64 (exact)
    o3 = torch.randn(37, 1012, 34, device=dev)
    t = nn_ops.obs_to_nhwc(o3, 1024)
    assert t.shape == (37, 1024, 1, 34) and t.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(t[:, :1012, 0, :], o3.to(torch.bfloat16)) and (t[:, 1012:] == 0).all()

    brain = Brain(conv_channels=192, num_blocks=6).to(dev).eval()
    for m in brain.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    obs = (torch.rand(64, 1012, 34, device=dev) < 0.05).float()
    with torch.inference_mode():
        ref = brain(obs)
        brain.prepare_fast(torch.bfloat16)
        fast = brain.forward_fast(obs).float()
    err = (fast - ref).abs().max().item() / ref.abs().max().item()
    assert err < 0.05, err  # bf16 end to end over 6 residual blocks


def test_device_engine_graph_replay_equals_eager(mjx):
    """DeviceEngine.react_static (CUDA-graph replay over persistent buffers) returns what react_device returns."""
    import torch

    from mortal_b200.engine import DeviceEngine
    from mortal_b200.model import DQN, Brain

    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    eng = DeviceEngine(Brain(conv_channels=64, num_blocks=3), DQN(), device=dev)
    obs = (torch.rand(700, 1012, 34, device=dev) < 0.05).float()
    masks = torch.rand(700, 46, device=dev) > 0.5
    masks[:, 45] = True
    for nr in (300, 511, 512, 650, 300):
        nb = min(-(-nr // 256) * 256, 700)  # the graph runs the padded batch; same shapes -> same kernels -> same bits
        a0, q0 = eng.react_device(obs[:nb], masks[:nb])
        a1, q1 = eng.react_static(obs, masks, nr)
        assert torch.equal(a0[:nr], a1) and torch.equal(q0[:nr], q1), nr
        # and against the un-padded eager call up to bf16 kernel-selection noise
        a2, q2 = eng.react_device(obs[:nr], masks[:nr])
        fin = torch.isfinite(q2)
        assert (q2[fin] - q1[fin]).abs().max() <= 0.05 * q2[fin].abs().max() + 1e-3
    assert len(eng._graphs) == 2  # buckets 512 and 700 (capped at the buffer size)


def test_gameplay_loader_on_device_matches_oracle(mjx, tmp_path):
    """libriichi.dataset.GameplayLoader on device — mjai logs written by the arena, loaded back through the
    device log replay, equal the oracle's restatement of dataset/gameplay.rs move for move (labels, masks, at_kyoku, dones,
    apply_gamma, at_turns, shantens) and observation for observation (v4 incl. the single-player block: exact)."""
    import gzip

    import torch

    from mortal_b200 import dataset_codec as DC
    from mortal_b200.libriichi.arena import OneVsThree
    from mortal_b200.libriichi.dataset import GameplayLoader

    class Rand:
        engine_type = "mortal"
        version = 4
        is_oracle = False
        enable_quick_eval = True
        enable_rule_based_agari_guard = False

        def __init__(self, name):
            self.name = name

        def react_device(self, obs, masks):
            # prefers shanten-lowering discards so that riichi / calls / kans all occur
            q = torch.rand(masks.shape, device=masks.device)
            q[:, :34] += 2.0 * obs[:, 876, :] + obs[:, 875, :]
            q = q.masked_fill(~masks, -1.0)
            return q.argmax(-1), q

    torch.manual_seed(5)
    arena = OneVsThree(disable_progress_bar=True, log_dir=str(tmp_path))
    arena.py_vs_py(Rand("x"), Rand("y"), (7000, 21), 2)
    files = sorted(str(p) for p in tmp_path.iterdir())
    assert len(files) == 8
    loader = GameplayLoader(4, oracle=False)
    loaded = loader.load_gz_log_files(files)
    assert len(loaded) == 8 and all(len(g) == 4 for g in loaded)
    moves = 0
    for fn, per_player in zip(files, loaded):
        events = DC.parse_log(gzip.open(fn, "rt").read())
        for gp in per_player:
            ref = O.gameplay_load(events, gp.take_player_id(), version=4, sp_mode=1)
            assert gp.take_actions() == ref["actions"].tolist()
            assert gp.take_at_kyoku() == ref["at_kyoku"].tolist() and gp.take_at_turns() == ref["at_turns"].tolist()
            assert gp.take_shantens() == ref["shantens"].tolist() and gp.take_apply_gamma() == ref["apply_gamma"].tolist()
            dones = np.append(ref["at_kyoku"][1:] > ref["at_kyoku"][:-1], True)
            assert gp.take_dones() == dones.tolist()
            assert (gp.take_masks(host=True) == ref["masks"]).all()
            obs = gp.take_obs(host=True)
            d = np.abs(obs - ref["obs"])
            assert obs.shape == ref["obs"].shape and not ((d != 0) & ((ref["obs"] == 0) | (ref["obs"] == 1))).any() and d.max() <= 1e-6
            assert (d[:, 889:] == 0).all()  # single-player block: exact
            assert gp.player_name == ("x" if gp.take_player_id() == files.index(fn) % 4 else "y")
            moves += len(ref["actions"])
    assert moves > 8 * 4 * 100
    only_x = GameplayLoader(4, oracle=False, player_names=["x"]).load_gz_log_files(files[:2])
    assert [len(g) for g in only_x] == [1, 1] and only_x[0][0].player_name == "x"
    # oracle=True (the reference's default): the invisible observation of every move. trust_seed: walls regenerated from
    # start_game.seed on device; otherwise reconstructed from the log with a random fill (the same walls handed to the oracle)
    orc = GameplayLoader(4, trust_seed=True).load_gz_log_files(files[:3])
    for fn, per_player in zip(files[:3], orc):
        events = DC.parse_log(gzip.open(fn, "rt").read())
        for gp in per_player:
            ref = O.gameplay_load(events, gp.take_player_id(), version=4, sp_mode=0, with_obs=False, oracle_seed=tuple(events[0]["seed"]))
            inv = gp.take_invisible_obs(host=True)
            assert inv.shape == ref["invisible"].shape and (inv == ref["invisible"]).all()
            assert gp.take_actions() == ref["actions"].tolist()
    rng = np.random.default_rng(11)
    loader2 = GameplayLoader(4, trust_seed=False, rng=rng)
    got2 = loader2.load_gz_log_files(files[:2])
    rng_ref = np.random.default_rng(11)
    for fn, per_player in zip(files[:2], got2):
        events = DC.parse_log(gzip.open(fn, "rt").read())
        walls = DC.reconstruct_walls(events, rng_ref)
        for gp in per_player:
            ref = O.gameplay_load(events, gp.take_player_id(), version=4, sp_mode=0, with_obs=False, walls=walls)
            assert (gp.take_invisible_obs(host=True) == ref["invisible"]).all()


def test_arena_agents_with_different_obs_version_and_quick_eval_on_device(mjx):
    """agent/mortal.rs:54-74, 256-287 on the CUDA path: a version-4 quick-eval challenger against a version-3 champion without
    quick-eval (device engines). Each engine is handed observations of its own layout; the recorded decisions and legal masks
    replay in the oracle under the same per-seat settings, to the same scores."""
    import torch

    from mortal_b200.libriichi.arena import OneVsThree

    class Eng:
        engine_type = "mortal"; is_oracle = False; enable_rule_based_agari_guard = False

        def __init__(self, name, version, qe):
            self.name, self.version, self.enable_quick_eval, self.rows = name, version, qe, 0

        def react_device(self, obs, masks):
            assert obs.shape[1:] == ({3: 934, 4: 1012}[self.version], 34) and float(obs.min()) >= 0.0 and float(obs.max()) <= 1.0
            self.rows += obs.shape[0]
            q = torch.rand(masks.shape, device=masks.device).masked_fill(~masks, -1.0)
            return q.argmax(-1), q

    torch.manual_seed(7)
    chal, champ = Eng("v4", 4, True), Eng("v3", 3, False)
    arena = OneVsThree(disable_progress_bar=True)
    arena.record_decisions = True
    rankings = arena.py_vs_py(chal, champ, (6600, 12), 4)
    assert sum(rankings) == 16 and chal.rows > 500 and champ.rows > 3 * chal.rows * 0.8
    n = 16
    nonces = np.repeat(np.arange(6600, 6604, dtype=np.uint64), 4)
    keys = np.full(n, 12, dtype=np.uint64)
    qf = np.array([[1 if seat == g % 4 else 0 for seat in range(4)] for g in range(n)], dtype=np.uint8)
    ref = O.run_replay(nonces, keys, arena.last_decisions, mask_bits=arena.last_decision_masks, quick_eval_seats=qf)
    assert (ref["scores"] == arena.last_results["scores"]).all() and (ref["ranks"] == arena.last_results["ranks"]).all()
