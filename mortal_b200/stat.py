"""libriichi.stat.Stat — per-player statistics over mjai logs (stat.rs:26-124 fields, 263-442 accumulation, 447-515 loaders,
516-786 derived rates). Host-side log arithmetic, no device work; used by mortal/player.py:71 and the test-play report of
mortal/train.py:321-350. Counters are plain integer attributes with the reference's names; every derived quantity is a
property `numerator / denominator` in float64 with IEEE semantics (0/0 = nan, x/0 = ±inf) like the Rust `as f64` divisions.
"""
from __future__ import annotations

import glob
import gzip
import json
import os

COUNTERS = (
    "game round oya point rank_1 rank_2 rank_3 rank_4 tobi "
    "agari agari_as_oya agari_jun agari_point_oya agari_point_ko "
    "riichi riichi_as_oya riichi_jun chasing_riichi riichi_got_chased riichi_agari riichi_agari_jun riichi_agari_point "
    "riichi_houjuu riichi_ryukyoku riichi_point "
    "fuuro fuuro_num fuuro_agari fuuro_agari_jun fuuro_agari_point fuuro_houjuu fuuro_point "
    "dama_agari dama_agari_jun dama_agari_point "
    "houjuu houjuu_jun houjuu_to_oya houjuu_point_to_oya houjuu_point_to_ko "
    "ryukyoku ryukyoku_point yakuman nagashi_mangan"
).split()

# derived quantity -> (numerator expression, denominator expression) over the counters (stat.rs:555-786)
RATES = {
    "rank_1_rate": ("rank_1", "game"), "rank_2_rate": ("rank_2", "game"), "rank_3_rate": ("rank_3", "game"),
    "rank_4_rate": ("rank_4", "game"), "tobi_rate": ("tobi", "game"),
    "avg_point_per_game": ("point", "game"), "avg_point_per_round": ("point", "round"),
    "avg_point_per_agari": ("agari_point_ko + agari_point_oya", "agari"),
    "avg_point_per_oya_agari": ("agari_point_oya", "agari_as_oya"),
    "avg_point_per_ko_agari": ("agari_point_ko", "agari - agari_as_oya"),
    "avg_point_per_riichi_agari": ("riichi_agari_point", "riichi_agari"),
    "avg_point_per_fuuro_agari": ("fuuro_agari_point", "fuuro_agari"),
    "avg_point_per_dama_agari": ("dama_agari_point", "dama_agari"),
    "avg_point_per_ryukyoku": ("ryukyoku_point", "ryukyoku"),
    "avg_agari_jun": ("agari_jun", "agari"), "avg_riichi_agari_jun": ("riichi_agari_jun", "riichi_agari"),
    "avg_fuuro_agari_jun": ("fuuro_agari_jun", "fuuro_agari"), "avg_dama_agari_jun": ("dama_agari_jun", "dama_agari"),
    "avg_point_per_houjuu": ("houjuu_point_to_ko + houjuu_point_to_oya", "houjuu"),
    "avg_point_per_houjuu_to_oya": ("houjuu_point_to_oya", "houjuu_to_oya"),
    "avg_point_per_houjuu_to_ko": ("houjuu_point_to_ko", "houjuu - houjuu_to_oya"),
    "avg_houjuu_jun": ("houjuu_jun", "houjuu"),
    "agari_rate": ("agari", "round"), "houjuu_rate": ("houjuu", "round"), "riichi_rate": ("riichi", "round"),
    "fuuro_rate": ("fuuro", "round"), "ryukyoku_rate": ("ryukyoku", "round"),
    "agari_rate_after_riichi": ("riichi_agari", "riichi"), "houjuu_rate_after_riichi": ("riichi_houjuu", "riichi"),
    "chasing_riichi_rate": ("chasing_riichi", "riichi"), "riichi_chased_rate": ("riichi_got_chased", "riichi"),
    "avg_riichi_jun": ("riichi_jun", "riichi"), "avg_riichi_point": ("riichi_point", "riichi"),
    "agari_rate_as_oya": ("agari_as_oya", "oya"), "agari_as_oya_rate": ("agari_as_oya", "agari"),
    "houjuu_to_oya_rate": ("houjuu_to_oya", "houjuu"),
    "avg_fuuro_num": ("fuuro_num", "fuuro"), "agari_rate_after_fuuro": ("fuuro_agari", "fuuro"),
    "houjuu_rate_after_fuuro": ("fuuro_houjuu", "fuuro"), "avg_fuuro_point": ("fuuro_point", "fuuro"),
    "yakuman_rate": ("yakuman", "round"), "nagashi_mangan_rate": ("nagashi_mangan", "round"),
}


def _fdiv(a: float, b: float) -> float:
    if b == 0:
        return float("nan") if a == 0 else (float("inf") if a > 0 else float("-inf"))
    return a / b


class Stat:
    def __init__(self, **counters):
        for name in COUNTERS:
            setattr(self, name, int(counters.get(name, 0)))

    def __getattr__(self, name):  # derived rates (only reached for names that are not counters)
        spec = RATES.get(name)
        if spec is None:
            raise AttributeError(name)
        env = {c: getattr(self, c) for c in COUNTERS}
        return _fdiv(float(eval(spec[0], {}, env)), float(eval(spec[1], {}, env)))

    def __add__(self, other: "Stat") -> "Stat":  # stat.rs Sum / Add: field-wise
        return Stat(**{c: getattr(self, c) + getattr(other, c) for c in COUNTERS})

    def __radd__(self, other):
        return self if other == 0 else self.__add__(other)

    def total_pt(self, pts) -> int:
        return self.rank_1 * pts[0] + self.rank_2 * pts[1] + self.rank_3 * pts[2] + self.rank_4 * pts[3]

    def avg_pt(self, pts) -> float:
        return _fdiv(float(self.total_pt(pts)), float(self.game))

    @property
    def avg_rank(self) -> float:
        return self.avg_pt([1, 2, 3, 4])

    # ---------------------------------------------------------------- stat.rs:263-442
    @staticmethod
    def from_game(events, player_id: int) -> "Stat":
        st = Stat(game=1)
        me = player_id
        scores = [0, 0, 0, 0]
        declared = accepted = others_declared = False
        oya = jun = calls = 0
        for ev in events:
            ty = ev["type"]
            if ty == "start_kyoku":
                st.round += 1
                scores = list(ev["scores"])
                declared = accepted = others_declared = False
                oya, jun, calls = ev["oya"], 0, 0
                st.oya += oya == me
            elif ty == "dahai":
                jun += ev["actor"] == me
            elif ty in ("chi", "pon", "daiminkan"):
                calls += ev["actor"] == me
            elif ty == "reach":
                if ev["actor"] == me:
                    declared = True
                    st.riichi += 1
                    st.riichi_jun += jun
                    st.riichi_as_oya += oya == me
                    st.chasing_riichi += others_declared
                elif declared:
                    st.riichi_got_chased += 1
                else:
                    others_declared = True
            elif ty == "reach_accepted":
                scores[ev["actor"]] -= 1000
                accepted = accepted or ev["actor"] == me
            elif ty == "hora":
                deltas = ev["deltas"]
                scores = [a + b for a, b in zip(scores, deltas)]
                if ev["actor"] == me:
                    point = deltas[me] - 1000 * accepted  # the own stick comes back with the win and is not counted
                    st.agari += 1
                    st.agari_jun += jun
                    if oya == me:
                        st.agari_as_oya += 1
                        st.agari_point_oya += point
                    else:
                        st.agari_point_ko += point
                    if accepted:
                        st.riichi_agari += 1; st.riichi_agari_jun += jun; st.riichi_agari_point += point; st.riichi_point += point
                    elif calls > 0:
                        st.fuuro_agari += 1; st.fuuro_agari_jun += jun; st.fuuro_agari_point += point; st.fuuro_point += point
                    else:
                        st.dama_agari += 1; st.dama_agari_jun += jun; st.dama_agari_point += point
                    st.yakuman += point >= (48000 if oya == me else 32000)  # point.rs Point::yakuman(is_oya, 1).ron
                elif ev["target"] == me:
                    point = deltas[me]
                    st.houjuu += 1
                    st.houjuu_jun += jun
                    if oya == ev["actor"]:
                        st.houjuu_to_oya += 1
                        st.houjuu_point_to_oya += point
                    else:
                        st.houjuu_point_to_ko += point
                    if declared:
                        st.riichi_houjuu += 1; st.riichi_point += point
                    elif calls > 0:
                        st.fuuro_houjuu += 1; st.fuuro_point += point
            elif ty == "ryukyoku":
                deltas = ev["deltas"]
                scores = [a + b for a, b in zip(scores, deltas)]
                point = deltas[me]
                st.ryukyoku += 1
                st.ryukyoku_point += point
                if accepted:
                    st.riichi_ryukyoku += 1
                    st.riichi_point += point - 1000
                elif calls > 0:
                    st.fuuro_point += point
                st.nagashi_mangan += point >= 8000
            elif ty == "end_kyoku":
                if calls > 0:
                    st.fuuro += 1
                    st.fuuro_num += calls
        order = sorted(range(4), key=lambda i: -scores[i])  # rankings.rs:8-22, stable by seat
        total = sum(scores)
        if total < 100_000:  # sticks left on the table go to the top
            scores[order[0]] += 100_000 - total
        st.point = scores[me] - 25000
        st.tobi = int(scores[me] < 0)
        setattr(st, ("rank_1", "rank_2", "rank_3", "rank_4")[order.index(me)], 1)
        return st

    @staticmethod
    def from_log(log: str, player_id: int) -> "Stat":
        return Stat.from_game([json.loads(ln) for ln in log.splitlines() if ln.strip()], player_id)

    @staticmethod
    def from_dir(dir: str, player_name: str, disable_progress_bar: bool = False) -> "Stat":
        total = Stat()
        paths = glob.glob(os.path.join(dir, "**", "*.json"), recursive=True) + glob.glob(os.path.join(dir, "**", "*.json.gz"), recursive=True)
        for path in paths:
            opener = gzip.open if path.lower().endswith(".gz") else open
            with opener(path, "rt") as f:
                events = [json.loads(ln) for ln in f if ln.strip()]
            if not events or events[0].get("type") != "start_game":
                raise ValueError(f"first event is not start_game, got {events[0] if events else None!r}")
            for i, name in enumerate(events[0].get("names", [])):
                if name == player_name:
                    total = total + Stat.from_game(events, i)
        return total

    def __repr__(self):
        return "Stat(" + ", ".join(f"{c}={getattr(self, c)}" for c in COUNTERS) + ")"

    def __str__(self):
        lines = [f"Games {self.game}", f"Rounds {self.round}", f"Rounds as dealer {self.oya}", ""]
        for k in (1, 2, 3, 4):
            lines.append(f"{k}{('st', 'nd', 'rd', 'th')[k - 1]} (rate) {getattr(self, f'rank_{k}')} ({getattr(self, f'rank_{k}_rate'):.6f})")
        lines += [f"Tobi(rate) {self.tobi} ({self.tobi_rate:.6f})", f"Avg rank {self.avg_rank:.6f}",
                  f"Total rank pt {self.total_pt([90, 45, 0, -135])}", f"Avg rank pt {self.avg_pt([90, 45, 0, -135]):.6f}",
                  f"Total score delta {self.point}", ""]
        lines += [f"{name} {getattr(self, name):.6f}" for name in RATES if not name.startswith("rank_") and name != "tobi_rate"]
        return "\n".join(lines)
