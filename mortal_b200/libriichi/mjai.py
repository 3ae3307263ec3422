"""libriichi.mjai — Bot (mjai/bot.rs:10-80): a single-seat mjai player over `libriichi.state.PlayerState` and a Mortal engine.

`Bot(engine, player_id).react(line, *, can_act=True) -> Optional[str]`: feed one mjai event (JSON string, optionally carrying
`can_act`), get the reaction event as JSON (with the `meta` agent/mortal.rs:161-186 attaches) or None. The decision logic is
MortalBatchAgent's for one seat (agent/mortal.rs:200-592): quick-eval shortcut, kan-select row queued before the normal row,
engine.react_batch over lists of arrays, rule-based agari guard, action id -> event decode (on device, mjx_state_query).
"""
from __future__ import annotations

import json
import time

import numpy as np

from ..mjai_log import TILE_NAMES, make_meta
from .state import PlayerState


class Bot:
    def __init__(self, engine, player_id: int):
        if getattr(engine, "engine_type", None) != "mortal":  # py_agent.rs:12-31: only `mortal` engines back a Bot
            raise ValueError(f"unknown engine type {getattr(engine, 'engine_type', None)!r}")
        if getattr(engine, "is_oracle", False):
            raise ValueError("an oracle engine needs the invisible state, which a Bot never has (mortal.rs:253-255)")
        self.engine = engine
        self.version = int(getattr(engine, "version", 4))
        self.enable_quick_eval = bool(getattr(engine, "enable_quick_eval", True))
        self.enable_rule_based_agari_guard = bool(getattr(engine, "enable_rule_based_agari_guard", False))
        self.player_id = int(player_id)
        self.state = PlayerState(self.player_id)

    def react(self, line: str, /, *, can_act: bool = True):
        try:
            data = json.loads(line)
        except Exception as exc:
            raise ValueError(f"failed to parse event {line}") from exc
        flag = data.pop("can_act", None)
        cans = self.state.update(data)
        if not can_act or flag is False or not cans.can_act:
            return None
        return json.dumps(self._decide(cans), separators=(",", ":"))

    # agent/mortal.rs:200-290 set_scene + 292-592 get_reaction for one seat
    def _decide(self, cans):
        st = self.state
        if (self.enable_quick_eval and cans.can_discard and not (cans.can_riichi or cans.can_tsumo_agari or cans.can_ankan
                                                                 or cans.can_kakan or cans.can_ryukyoku)):
            cand = np.nonzero(st.discard_candidates())[0]
            if len(cand) == 1:  # mortal.rs:210-242: the only legal discard is played without asking the engine; no meta
                pai = TILE_NAMES[int(cand[0])]
                return {"type": "dahai", "actor": self.player_id, "pai": pai, "tsumogiri": st.last_self_tsumo() == pai}
        need_kan = (cans.can_ankan or cans.can_kakan) and (not self.enable_quick_eval or
                                                           len(st.ankan_candidates()) + len(st.kakan_candidates()) > 1)
        obs_list, mask_list = [], []
        if need_kan:
            o, m = st.encode_obs(self.version, True)
            obs_list.append(o); mask_list.append(m)
        o, m = st.encode_obs(self.version, False)
        obs_list.append(o); mask_list.append(m)
        t0 = time.perf_counter_ns()
        actions, q_values, masks, is_greedy = self.engine.react_batch(obs_list, mask_list, None)
        eval_ns = time.perf_counter_ns() - t0
        idx = len(obs_list) - 1
        action = int(actions[idx])
        q = np.asarray(q_values[idx], dtype=np.float32)
        legal = np.asarray(masks[idx], dtype=bool)
        if self.enable_rule_based_agari_guard and action == 43 and not st.rule_based_agari():  # mortal.rs:319-336
            q2 = q.copy()
            q2[43] = np.finfo(np.float32).min
            best = 0
            for i in range(46):  # max_by(total_cmp): the LAST maximum
                if not (q2[i] < q2[best]):
                    best = i
            action = best
        kan_action = int(actions[0]) if need_kan else -1
        reaction = st.decode_action(action, kan_action if action == 42 else -1)
        v = st.view()
        common = dict(batch_size=len(obs_list), eval_time_ns=int(eval_ns), shanten=int(v.shanten), at_furiten=bool(v.at_furiten))
        kan_meta = None
        if need_kan and action == 42:
            km = make_meta(kan_action, np.asarray(masks[0], dtype=bool), np.asarray(q_values[0], dtype=np.float32),
                           is_greedy=bool(is_greedy[0]), **common)
            kan_meta = {k: val for k, val in km.items() if not k.startswith("_") and val is not None}
        meta = make_meta(action, legal, q, is_greedy=bool(is_greedy[idx]), kan_select=kan_meta, **common)
        reaction["meta"] = {k: val for k, val in meta.items() if not k.startswith("_") and val is not None}
        return reaction
