"""mjai log emission: device event words -> the reference's JSON lines and .json.gz files.

Format contract: libriichi mjai/event.rs:20-120 (serde: internally tagged `type` first, fields in declaration order,
`None` options skipped), arena/result.rs:32-51 (`start_game` with names + seed, every kyoku's events, `end_game`),
file names arena/one_vs_three.rs:203-216 / two_vs_two.rs:213-215 (`{seed}_{key}_{a|b|c|d}.json.gz`).
The words are produced on device by csrc/mjx_step.cuh (`log_word`); per-decision `meta` is not recorded.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

TILE_NAMES = ([f"{n}{s}" for s in "mps" for n in range(1, 10)] + ["E", "S", "W", "N", "P", "F", "C"]
              + ["5mr", "5pr", "5sr", "?"])  # tile.rs:12-19 ids 0..37

(START_KYOKU, TSUMO, DAHAI, CHI, PON, DAIMINKAN, KAKAN, ANKAN, DORA, REACH, REACH_ACCEPTED, HORA, RYUKYOKU,
 END_KYOKU) = range(1, 15)


def _i32x4(w0: int, w1: int):
    v = [w0 & 0xFFFFFFFF, w0 >> 32, w1 & 0xFFFFFFFF, w1 >> 32]
    return [x - (1 << 32) if x >= (1 << 31) else x for x in v]


def decode_events(words, with_offsets: bool = False):
    """One table's words -> list of mjai event dicts (field order as serde writes them); with_offsets also returns the
    word index each event starts at."""
    out, offs = [], []
    i, n = 0, len(words)
    while i < n:
        w = int(words[i])
        offs.append(i)
        i += 1
        ty = w & 0xFF
        actor, target = (w >> 8) & 3, (w >> 10) & 3
        pai = (w >> 12) & 0xFF
        tsumogiri = bool((w >> 20) & 1)
        aux = (w >> 21) & 7
        c = [(w >> s) & 0xFF for s in (24, 32, 40, 48)]
        extra = (w >> 56) & 0xFF
        t = TILE_NAMES
        if ty == START_KYOKU:
            kyoku_abs, honba, kyotaku, oya = c
            scores = _i32x4(int(words[i]), int(words[i + 1]))
            raw = b"".join(int(x).to_bytes(8, "little") for x in words[i + 2:i + 9])[:52]
            i += 9
            tehais = [[t[raw[13 * s + k]] for k in range(13)] for s in range(4)]
            out.append({"type": "start_kyoku", "bakaze": t[27 + kyoku_abs // 4], "dora_marker": t[pai], "kyoku": oya + 1,
                        "honba": honba, "kyotaku": kyotaku, "oya": oya, "scores": scores, "tehais": tehais})
        elif ty == TSUMO:
            out.append({"type": "tsumo", "actor": actor, "pai": t[pai]})
        elif ty == DAHAI:
            out.append({"type": "dahai", "actor": actor, "pai": t[pai], "tsumogiri": tsumogiri})
        elif ty in (CHI, PON):
            out.append({"type": "chi" if ty == CHI else "pon", "actor": actor, "target": target, "pai": t[pai],
                        "consumed": [t[c[0]], t[c[1]]]})
        elif ty == DAIMINKAN:
            out.append({"type": "daiminkan", "actor": actor, "target": target, "pai": t[pai],
                        "consumed": [t[c[0]], t[c[1]], t[c[2]]]})
        elif ty == KAKAN:
            out.append({"type": "kakan", "actor": actor, "pai": t[pai], "consumed": [t[c[0]], t[c[1]], t[c[2]]]})
        elif ty == ANKAN:
            out.append({"type": "ankan", "actor": actor, "consumed": [t[x] for x in c]})
        elif ty == DORA:
            out.append({"type": "dora", "dora_marker": t[pai]})
        elif ty == REACH:
            out.append({"type": "reach", "actor": actor})
        elif ty == REACH_ACCEPTED:
            out.append({"type": "reach_accepted", "actor": actor})
        elif ty == HORA:
            deltas = _i32x4(int(words[i]), int(words[i + 1]))
            i += 2
            # ura_markers is always Some(..): empty unless the winner's riichi was accepted (board.rs:423-432)
            out.append({"type": "hora", "actor": actor, "target": target, "deltas": deltas,
                        "ura_markers": [t[x] for x in (c + [extra])[:aux]]})
        elif ty == RYUKYOKU:
            deltas = _i32x4(int(words[i]), int(words[i + 1]))
            i += 2
            out.append({"type": "ryukyoku", "deltas": deltas})
        elif ty == END_KYOKU:
            out.append({"type": "end_kyoku"})
        else:
            raise ValueError(f"corrupt event log: word {w:#x} at {i - 1}")
    return (out, offs) if with_offsets else out


AGENT_EVENT_ACTIONS = {  # event type -> action ids that produce it (mortal.rs:292-573 decode)
    "dahai": range(0, 37), "reach": (37,), "chi": (38, 39, 40), "pon": (41,), "daiminkan": (42,), "kakan": (42,), "ankan": (42,),
    "hora": (43,), "ryukyoku": (44,),
}


def attach_meta(events: list[dict], offsets, bounds, decisions: dict) -> int:
    """Attach the per-decision `meta` (mjai/event.rs:131-150 Metadata, filled as agent/mortal.rs:161-186 gen_meta fills it) to the
    agent events of one game, in place.

    offsets[i]  word index at which events[i] starts in the table's device log (decode_events(..., with_offsets=True)),
    bounds[c]   the table's log length right after environment step c,
    decisions   c -> {seat: meta dict with an extra "_action" key}: the decisions taken on the rows step c emitted.
    An event written during step c was caused by a decision of step c - 1; it gets that seat's meta when the decision's action
    produces this event type. Quick-evaluated decisions have no row and, as in the reference, no meta.
    Returns the number of events that received a meta."""
    import bisect

    attached = 0
    for ev, off in zip(events, offsets):
        acts = AGENT_EVENT_ACTIONS.get(ev["type"])
        if acts is None:
            continue
        cand = decisions.get(bisect.bisect_right(bounds, off) - 1)
        if not cand:
            continue
        if ev["type"] == "ryukyoku":  # no actor field: the seat that declared it (abortive draws by rule have none)
            seats = [s for s, m in cand.items() if m["_action"] == 44]
        else:
            seats = [ev["actor"]] if ev["actor"] in cand and cand[ev["actor"]]["_action"] in acts else []
        if seats:
            ev["meta"] = {key: val for key, val in cand[seats[0]].items() if not key.startswith("_") and val is not None}
            attached += 1
    return attached


def make_meta(action: int, mask_row, q_row, *, is_greedy=True, batch_size=None, eval_time_ns=None, shanten=None, at_furiten=None,
              kan_select=None) -> dict:
    """Metadata in the reference's field order; q_values holds the legal actions only, in action order (mortal.rs:166-176)."""
    legal = [i for i in range(46) if mask_row[i]]
    return {"_action": int(action), "q_values": [float(q_row[i]) for i in legal], "mask_bits": sum(1 << i for i in legal),
            "is_greedy": bool(is_greedy), "batch_size": batch_size, "eval_time_ns": eval_time_ns, "shanten": shanten,
            "at_furiten": at_furiten, "kan_select": kan_select}


def dump_json_log(events: list[dict], names, seed) -> str:
    """arena/result.rs:32-51 GameResult::dump_json_log"""
    dumps = lambda o: json.dumps(o, separators=(",", ":"), ensure_ascii=False)
    lines = [dumps({"type": "start_game", "names": list(names), "seed": [int(seed[0]), int(seed[1])]})]
    lines += [dumps(e) for e in events]
    lines.append(dumps({"type": "end_game"}))
    return "\n".join(lines) + "\n"


def write_logs(log_dir: str, words: np.ndarray, lens: np.ndarray, seeds, names_per_game, split_names, bounds=None, decisions=None) -> list[str]:
    """one `{seed}_{key}_{split}.json.gz` per game (one_vs_three.rs:203-216); returns the paths.
    bounds [n_steps, n_games] / decisions[g] (see attach_meta) add the per-decision meta when given."""
    os.makedirs(log_dir, exist_ok=True)
    paths = []
    for g in range(words.shape[0]):
        events, offsets = decode_events(words[g, : int(lens[g])], with_offsets=True)
        if bounds is not None and decisions is not None:
            attach_meta(events, offsets, [int(b) for b in bounds[:, g]], decisions[g])
        text = dump_json_log(events, names_per_game[g], seeds[g])
        path = os.path.join(log_dir, f"{int(seeds[g][0])}_{int(seeds[g][1])}_{split_names[g % len(split_names)]}.json.gz")
        with gzip.open(path, "wb", compresslevel=9) as f:
            f.write(text.encode("utf-8"))
        paths.append(path)
    return paths
