"""Pins the oracle's single-player calculator to the reference KATs (algo/sp/calc.rs:772-1008, default features)."""
import ctypes as C

import numpy as np

import oracle_lib as O
from oracle_lib import SpCand, SpIn, hand, tid

EPS = float(np.finfo(np.float32).eps)


def feq(a, b):
    return abs(np.float32(a) - np.float32(b)) <= EPS


def run(tehai, *, jikaze, dora, tsumos_left, can_discard, seen_extra=None, akas_seen=(0, 0, 0), double_riichi=False,
        haitei=False, maximize_win_prob=False, tegawari=True, shanten_down=True):
    q = SpIn()
    h = hand(tehai)
    seen = h.copy()
    seen[tid(dora)] += 1
    if seen_extra:
        for k, v in seen_extra.items():
            seen[tid(k)] += v
    for i in range(34):
        q.tehai[i] = int(h[i])
        q.tiles_seen[i] = int(seen[i])
    for i in range(3):
        q.akas_seen[i] = akas_seen[i]
    q.tehai_len_div3 = 4
    q.is_menzen = 1
    q.bakaze = tid("E")
    q.jikaze = tid(jikaze)
    q.n_dora_indicators = 1
    q.dora_indicators[0] = tid(dora)
    q.calc_double_riichi = int(double_riichi)
    q.calc_haitei = int(haitei)
    q.prefer_riichi = 1
    q.sort_result = 1
    q.maximize_win_prob = int(maximize_win_prob)
    q.calc_tegawari = int(tegawari)
    q.calc_shanten_down = int(shanten_down)
    q.can_discard = int(can_discard)
    q.tsumos_left = tsumos_left
    q.cur_shanten = int(O.shanten(h, 4)[0])
    out = (SpCand * 16)()
    n = O.lib().orc_sp_calc(C.byref(q), out, 16)
    assert n >= 0, O.err()
    return [out[i] for i in range(n)], seen


def test_nanikiru():  # calc.rs:772-942
    c, _ = run("45678m 34789p 3344z", jikaze="N", dora="P", tsumos_left=8, can_discard=True)
    assert c[0].tile == tid("N") and c[1].tile == tid("W")
    assert list(c[0].exp_values[: c[0].n_turns]) > list(c[1].exp_values[: c[1].n_turns])

    c, _ = run("3667m 23489p 34688s", jikaze="N", dora="P", tsumos_left=15, can_discard=True)
    assert c[0].tile == tid("9p") and c[0].shanten_down
    c, _ = run("3667m 23489p 34688s", jikaze="N", dora="P", tsumos_left=15, can_discard=True, maximize_win_prob=True)
    assert c[0].tile == tid("3m") and not c[0].shanten_down

    c, _ = run("45677m 456778p 248s", jikaze="E", dora="6m", tsumos_left=15, can_discard=True, double_riichi=True,
               haitei=True)
    c0 = c[0]
    assert c0.tile == tid("2s") and c0.n_required == 17 and c0.num_required_tiles == 57 and c0.shanten_down
    assert feq(c0.tenpai_probs[0], 0.90023905) and feq(c0.win_probs[0], 0.34794784) and feq(c0.exp_values[0], 5894.7617)

    c, seen = run("9999m 6677p 88s 335z 1m", jikaze="W", dora="1m", tsumos_left=5, can_discard=True)
    assert len(c) == 7
    c1 = c[1]
    assert c1.tile == tid("1m") and c1.shanten_down and c1.n_required == 33
    assert c1.num_required_tiles == 34 * 4 - int(seen.sum())


def test_tsumo_only():  # calc.rs:944-1007
    c, _ = run("45677m 456778p 48s", jikaze="W", dora="6m", tsumos_left=5, can_discard=False, double_riichi=True,
               haitei=True, maximize_win_prob=True, seen_extra={"5s": 4}, akas_seen=(0, 0, 1))
    assert len(c) == 1
    c0 = c[0]
    assert c0.tile == 37 and c0.n_required == 16 and c0.num_required_tiles == 54
    assert feq(c0.tenpai_probs[0], 0.45017204) and feq(c0.win_probs[0], 0.03441279) and feq(c0.exp_values[0], 432.26678)
