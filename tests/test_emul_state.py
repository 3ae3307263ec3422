"""libriichi.state.PlayerState on the host-emulated product: the reference's state/test.rs assertions
(tests/state_cases.py) through mortal_b200.libriichi.state with the emulation backend, and the device observation of a
single-seat state against the oracle's PlayerState fed the same partial-information events."""
import json

import numpy as np
import pytest

import oracle_lib as O
import state_cases as SC


@pytest.fixture(scope="module")
def PS():
    from emul_state import EmulStateBackend
    from mortal_b200.libriichi import state

    state.set_backend(EmulStateBackend())
    yield state.PlayerState
    state.set_backend(None)


@pytest.mark.parametrize("case", SC.ALL_CASES, ids=lambda c: c.__name__)
def test_state_test_rs_cases(PS, case):
    case(PS)


def test_partial_information_obs_equals_oracle_player_state(PS):
    """Every log of state/test.rs (single-seat view: other hands are `?`) replayed through the device PlayerState and through the
    oracle's: the ActionCandidate of every update, and the full v4 observation + mask (incl. the single-player tables) wherever the
    seat can act, are identical."""
    n_obs = 0
    for name, logs in SC.LOGS.items():
        for pid in range(4):
            lines = logs[0]
            first = json.loads(lines[0])
            if first["type"] != "start_kyoku" or first["tehais"][pid][0] == "?":
                continue
            ps, ref = PS(pid), O.PlayerState(pid)
            for ln in lines:
                got, want = ps.update(ln), ref.update(ln)
                assert {k: got[k] for k in O.CAN_BITS} == {k: want[k] for k in O.CAN_BITS} and got.target_actor == want["target_actor"], (name, ln)
                if got.can_act:
                    obs, mask = ps.encode_obs(4, False)
                    robs, rmask = ref.encode_obs(4, False, sp_mode=1)
                    assert (mask == rmask).all(), (name, ln)
                    d = np.abs(obs - robs)
                    assert not ((d != 0) & ((robs == 0) | (robs == 1))).any() and d.max() <= 1e-6, (name, ln, np.argwhere(d > 1e-6)[:4])
                    assert (d[889:] == 0).all()
                    n_obs += 1
    assert n_obs > 100


def test_mjai_bot_plays_a_seat_of_a_logged_game(PS):
    """libriichi.mjai.Bot (mjai/bot.rs:10-80) fed the golden game from seat 1's point of view (other hands hidden): whenever the log
    shows that seat acting, the Bot — driven by an engine that always picks the logged action — must answer with exactly the logged
    event; its meta carries the reference-written mask_bits; and `can_act=False` only updates the state."""
    from mortal_b200.libriichi.mjai import Bot
    from test_oracle_golden import AGENT_EVENTS, load_golden, strip_meta

    golden = load_golden()
    pid = 1
    tile_id = {n: i for i, n in enumerate(O.TILE_NAMES)}

    def action_of(ev):
        if ev["type"] == "dahai":
            return tile_id[ev["pai"]]
        return {"reach": 37, "pon": 41, "hora": 43}.get(ev["type"], None)

    class Scripted:
        engine_type = "mortal"; version = 4; is_oracle = False; enable_quick_eval = False; enable_rule_based_agari_guard = False; name = "s"
        want = 45

        def react_batch(self, obs, masks, invisible_obs):
            assert invisible_obs is None and obs[0].shape == (1012, 34)
            m = np.stack(masks)
            a = [self.want if m[i, self.want] else int(np.nonzero(m[i])[0][-1]) for i in range(len(obs))]
            return a, np.where(m, 1.0, -np.inf).tolist(), m.tolist(), [True] * len(obs)

    eng = Scripted()
    bot = Bot(eng, pid)
    hidden = lambda e: ({**e, "tehais": [h if s == pid else ["?"] * 13 for s, h in enumerate(e["tehais"])]} if e["type"] == "start_kyoku"
                        else ({**e, "pai": "?"} if e["type"] == "tsumo" and e["actor"] != pid else e))
    answered = 0
    for i, raw in enumerate(golden):
        ev = hidden(strip_meta(raw))
        nxt = next((strip_meta(g) for g in golden[i + 1:i + 3] if g["type"] not in ("reach_accepted", "dora")), None)
        mine = nxt is not None and nxt.get("actor") == pid and nxt["type"] in AGENT_EVENTS | {"hora"} and action_of(nxt) is not None
        eng.want = action_of(nxt) if mine else 45
        out = bot.react(json.dumps(ev))
        if mine and out is not None:
            got = json.loads(out)
            meta = got.pop("meta")
            want = {k: v for k, v in nxt.items() if k not in ("deltas", "ura_markers")}
            assert got == want, (i, got, want)
            ref_meta = golden[i + 1].get("meta") or golden[min(i + 2, len(golden) - 1)].get("meta")
            if ref_meta and "mask_bits" in ref_meta and golden[i + 1].get("actor") == pid:
                assert meta["mask_bits"] == ref_meta["mask_bits"]
            assert len(meta["q_values"]) == bin(meta["mask_bits"]).count("1")
            answered += 1
    assert answered > 20
    b2 = Bot(eng, 0)
    assert b2.react(json.dumps(hidden(strip_meta(golden[1]))), can_act=False) is None
