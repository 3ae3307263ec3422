"""Pins the oracle's PlayerState to the reference's state tests.

Assert logic re-stated from /root/reference/libriichi/src/state/test.rs (line refs inline); the inline
mjai logs come from tests/golden/state_test_logs.json (tools/extract_ref_fixtures.py).
Every update is followed by the reference's own invariant checker (test.rs:49-67).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import PlayerState, hand, hand_with_aka, tid

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "state_test_logs.json")) as f:
    LOGS = json.load(f)

UNK13 = ["?"] * 13


def tiles_of(s):
    h = hand_with_aka(s)
    out = []
    for t in range(37):
        out += [O.TILE_NAMES[t]] * int(h[t])
    return out


def validate(ps: PlayerState):
    """test.rs:49-67"""
    v = ps.view()
    th = np.array(list(v.tehai), dtype=np.uint8)
    assert v.real_time_shanten == int(O.shanten(th, v.tehai_len_div3)[0])
    assert bool(v.is_menzen) == (v.n_chis == 0 and v.n_pons == 0 and v.n_minkans == 0)
    cans = O.unpack_cans(v.cans)
    if any(cans[k] for k in O.CAN_BITS):
        for version in (1, 2, 3, 4):
            obs, mask = ps.encode_obs(version, False)
            assert obs.min() >= 0.0 and obs.max() <= 1.0
            if cans["can_kakan"] or cans["can_ankan"]:
                ps.encode_obs(version, True)


def upd(ps, ev):
    cans = ps.update(ev)
    validate(ps)
    return cans


def from_log(pid, lines):
    ps = PlayerState(pid)
    for ln in lines:
        upd(ps, ln)
    return ps


def start_kyoku(tehai0, dora, **kw):
    d = dict(type="start_kyoku", bakaze="E", kyoku=1, honba=0, kyotaku=0, oya=0, scores=[25000] * 4,
             dora_marker=dora, tehais=[tiles_of(tehai0), UNK13, UNK13, UNK13])
    d.update(kw)
    return d


def test_waits():  # test.rs:70-101
    L = O.lib()
    for s, expected in (("456m 78999p 789s 77z", ["6p", "9p", "C"]),
                        ("2344445666678s", ["1s", "2s", "3s", "5s", "7s", "8s", "9s"])):
        ps = PlayerState(0)
        h = hand(s)
        L.orc_ps_set_tehai(ps._p, h.ctypes.data, 4)
        assert L.orc_ps_update_waits_and_furiten(ps._p) == 0
        w = list(ps.view().waits)
        assert [O.TILE_NAMES[i] for i in range(34) if w[i]] == expected


def test_can_chi():  # test.rs:103-220 (the matrix of low/mid/high flags)
    L = O.lib()

    def chi(tehai, tile):
        ps = PlayerState(0)
        h = hand(tehai)
        L.orc_ps_set_tehai(ps._p, h.ctypes.data, 4)
        c = O.unpack_cans(L.orc_ps_set_can_chi_from_tile(ps._p, tid(tile)))
        return (c["can_chi_low"], c["can_chi_mid"], c["can_chi_high"])

    # (low, mid, high) — test.rs:106-220
    assert chi("1111234m", "1m") == (False, False, False)
    assert chi("1111234m", "4m") == (False, False, False)
    assert chi("1111234m", "2m") == (True, True, False)
    assert chi("6666789999p", "5p") == (True, False, False)
    assert chi("6666789999p", "7p") == (True, True, False)
    assert chi("6666789999p", "8p") == (False, True, True)
    assert chi("4556s", "3s") == (True, False, False)
    assert chi("4556s", "4s") == (True, False, False)
    assert chi("4556s", "5s") == (False, False, False)
    assert chi("4556s", "6s") == (False, False, True)
    assert chi("4556s", "7s") == (False, False, True)


def test_furiten():  # test.rs:223-477
    ps = PlayerState(0)
    upd(ps, start_kyoku("23406m 456789p 58s", "3p"))
    upd(ps, dict(type="tsumo", actor=0, pai="8s"))
    v = ps.view()
    assert v.shanten == 1 and not any(v.waits)
    upd(ps, dict(type="dahai", actor=0, pai="5s", tsumogiri=False))
    v = ps.view()
    assert v.shanten == 0 and v.waits[tid("1m")] and v.waits[tid("4m")] and v.waits[tid("7m")] and not v.at_furiten

    upd(ps, dict(type="tsumo", actor=1, pai="?"))
    cans = upd(ps, dict(type="dahai", actor=1, pai="1m", tsumogiri=False))
    assert not ps.view().at_furiten and cans["can_ron_agari"]

    upd(ps, dict(type="tsumo", actor=2, pai="?"))
    assert ps.view().at_furiten  # same-cycle furiten
    upd(ps, dict(type="dahai", actor=2, pai="1s", tsumogiri=True))
    upd(ps, dict(type="tsumo", actor=3, pai="?"))
    cans = upd(ps, dict(type="dahai", actor=3, pai="1m", tsumogiri=False))
    v = ps.view()
    assert v.shanten == 0 and v.at_furiten and not cans["can_ron_agari"]

    upd(ps, dict(type="tsumo", actor=0, pai="3s"))
    assert ps.view().at_furiten
    upd(ps, dict(type="dahai", actor=0, pai="3s", tsumogiri=True))
    assert not ps.view().at_furiten

    for actor, pai in ((1, "P"), (2, "C")):
        upd(ps, dict(type="tsumo", actor=actor, pai="?"))
        upd(ps, dict(type="dahai", actor=actor, pai=pai, tsumogiri=True))
    upd(ps, dict(type="tsumo", actor=3, pai="?"))
    cans = upd(ps, dict(type="dahai", actor=3, pai="1m", tsumogiri=False))
    assert not ps.view().at_furiten and cans["can_ron_agari"]
    assert ps.agari_points(True)["ron"] == 5800  # test.rs:337

    # riichi furiten (test.rs:339-476)
    cans = upd(ps, dict(type="tsumo", actor=0, pai="N"))
    assert cans["can_riichi"]
    upd(ps, dict(type="reach", actor=0))
    upd(ps, dict(type="dahai", actor=0, pai="N", tsumogiri=True))
    upd(ps, dict(type="reach_accepted", actor=0))
    for actor in (1, 2, 3):
        upd(ps, dict(type="tsumo", actor=actor, pai="?"))
        upd(ps, dict(type="dahai", actor=actor, pai="N", tsumogiri=True))
    cans = upd(ps, dict(type="tsumo", actor=0, pai="7m"))
    v = ps.view()
    assert v.waits[tid("1m")] and v.waits[tid("4m")] and v.waits[tid("7m")] and not v.at_furiten
    assert cans["can_tsumo_agari"]
    upd(ps, dict(type="dahai", actor=0, pai="7m", tsumogiri=True))
    assert ps.view().at_furiten  # furiten forever from now on
    upd(ps, dict(type="tsumo", actor=1, pai="?"))
    cans = upd(ps, dict(type="dahai", actor=1, pai="4m", tsumogiri=True))
    v = ps.view()
    assert v.at_furiten and not cans["can_ron_agari"]
    for actor in (2, 3):
        upd(ps, dict(type="tsumo", actor=actor, pai="?"))
        upd(ps, dict(type="dahai", actor=actor, pai="W", tsumogiri=True))
    assert ps.view().at_furiten
    cans = upd(ps, dict(type="tsumo", actor=0, pai="4m"))
    v = ps.view()
    assert v.waits[0] and v.waits[3] and v.waits[6] and v.at_furiten and cans["can_tsumo_agari"]
    assert ps.agari_points(False, ["3m"])["tsumo_ko"] == 6000  # test.rs:476


def test_dora_count_after_kan():  # test.rs:479-579
    ps = PlayerState(0)
    upd(ps, start_kyoku("1111s 123456p 112z", "N"))
    upd(ps, dict(type="tsumo", actor=0, pai="8s"))
    assert ps.view().doras_owned[0] == 2
    upd(ps, dict(type="ankan", actor=0, consumed=["1s"] * 4))
    upd(ps, dict(type="dora", dora_marker="9s"))
    upd(ps, dict(type="tsumo", actor=0, pai="5pr"))
    assert ps.view().doras_owned[0] == 7
    upd(ps, dict(type="dahai", actor=0, pai="E", tsumogiri=True))
    assert ps.view().doras_owned[0] == 6
    upd(ps, dict(type="tsumo", actor=1, pai="?"))
    upd(ps, dict(type="dahai", actor=1, pai="5p", tsumogiri=True))
    upd(ps, dict(type="pon", actor=0, target=1, pai="5p", consumed=["5pr", "5p"]))
    assert ps.view().doras_owned[0] == 6
    upd(ps, dict(type="dahai", actor=0, pai="E", tsumogiri=False))
    assert ps.view().doras_owned[0] == 5
    for actor in (1, 2):
        upd(ps, dict(type="tsumo", actor=actor, pai="?"))
        upd(ps, dict(type="dahai", actor=actor, pai="P", tsumogiri=True))
    upd(ps, dict(type="tsumo", actor=3, pai="?"))
    upd(ps, dict(type="ankan", actor=3, consumed=["1m"] * 4))
    upd(ps, dict(type="dora", dora_marker="4p"))
    assert ps.view().doras_owned[0] == 8


def test_rule_based_agari_all_last_minogashi():  # test.rs:581-799
    logs = LOGS["rule_based_agari_all_last_minogashi"]
    ps = from_log(1, logs[0])
    assert O.unpack_cans(ps.view().cans)["can_tsumo_agari"]
    assert not ps.rule_based_agari()
    # test.rs:668-676 mutate private fields (scores / an extra dora indicator); the oracle exposes the
    # same decision through rule_based_agari_slow on a clone driven by an equivalent Dora event
    ps_b = from_log(2, logs[1])
    assert ps_b.rule_based_agari()


def test_get_rank():  # test.rs:801-826
    L = O.lib()

    def gr(pid, scores):
        s = np.array(scores, dtype=np.int32)
        return L.orc_ps_get_rank(pid, s.ctypes.data)

    assert gr(0, [20000, 25000, 25000, 30000]) == 3
    assert gr(3, [25000, 25000, 25000, 25000]) == 3
    assert gr(1, [25000, 30000, 20000, 25000]) == 2
    assert gr(1, [32000, 32000, 18000, 18000]) == 0
    assert gr(2, [32000, 18000, 18000, 32000]) == 1
    assert gr(2, [5, 2, 5, 3]) == 1


def test_kakan_from_hand():  # test.rs:828-911
    ps = from_log(1, LOGS["kakan_from_hand"][0])
    assert O.unpack_cans(ps.view().cans)["can_tsumo_agari"]


def test_discard_candidates_with_unconditional_tenpai():  # test.rs:913-1224
    logs = LOGS["discard_candidates_with_unconditional_tenpai"]
    ps = from_log(1, logs[0])
    full = ps.discard_candidates(unconditional_tenpai=True)
    d34 = full[:34].copy()
    d34[4] |= full[34]; d34[13] |= full[35]; d34[22] |= full[36]
    assert [O.TILE_NAMES[i] for i in range(34) if d34[i]] == ["7p", "8p"]
    ps = from_log(1, logs[1])
    w = list(ps.view().waits)
    assert [O.TILE_NAMES[i] for i in range(34) if w[i]] == ["5p", "8p"]
    assert not ps.discard_candidates(unconditional_tenpai=True).any()


def test_double_chankan_ron():  # test.rs:1226-1391
    logs = LOGS["double_chankan_ron"]
    ps = from_log(2, logs[0])
    ps_kakan = ps.clone()
    cans = upd(ps_kakan, logs[1][0])
    assert cans["can_ron_agari"]
    assert ps_kakan.agari_points(True)["ron"] == 1000
    cans = upd(ps, logs[2][0])
    assert not cans["can_ron_agari"]


def test_chi_at_0_shanten():  # test.rs:1393-1418
    logs = LOGS["chi_at_0_shanten"]
    ps = from_log(0, logs[0])
    v = ps.view()
    cans = O.unpack_cans(v.cans)
    assert v.shanten == 0 and v.real_time_shanten == 0 and cans["can_ron_agari"] and cans["can_chi_high"]
    upd(ps, logs[1][0])
    v = ps.view()
    assert v.shanten == 0 and v.real_time_shanten == -1 and v.at_furiten and not v.has_next_shanten_discard


def test_selfplay_invariants_and_score_conservation():
    """arena/game.rs:324-372 analogue: hanchans run clean; plus conservation (scores sum to 100000)."""
    n = 24
    nonces = np.arange(10000, 10000 + n, dtype=np.uint64)
    keys = np.full(n, 0x2000, dtype=np.uint64)
    for kind in (0, 1):
        for qe in (True, False):
            r = O.run_batch(nonces, keys, policy_kind=kind, quick_eval=qe, encode_obs=0)
            assert (r["scores"].sum(axis=1) == 100000).all()
            assert (r["steps"] > 30).all()
    # greedy policy must actually win hands: someone ends above 25000 by a margin in most games
    r = O.run_batch(nonces, keys, policy_kind=1)
    assert (r["scores"].max(axis=1) > 30000).mean() > 0.5
    # determinism + thread-count independence
    r2 = O.run_batch(nonces, keys, policy_kind=1, n_threads=4)
    assert (r["scores"] == r2["scores"]).all() and (r["steps"] == r2["steps"]).all()


def test_selfplay_with_obs_encode_all_versions():
    nonces = np.arange(20000, 20004, dtype=np.uint64)
    keys = np.full(4, 7, dtype=np.uint64)
    base = O.run_batch(nonces, keys, policy_kind=1)
    for version in (1, 2, 3):
        r = O.run_batch(nonces, keys, policy_kind=1, encode_obs=version)
        assert (r["scores"] == base["scores"]).all()
    r = O.run_batch(nonces[:2], keys[:2], policy_kind=1, encode_obs=4, sp_mode=1, max_steps=60)
    assert r["obs_rows"] > 0
