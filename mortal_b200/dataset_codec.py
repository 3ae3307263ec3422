"""mjai events -> the compact event words the replay kernel consumes (the inverse of mortal_b200.mjai_log.decode_events;
word layout in csrc/mjx_step.cuh `log_word`, job layout in csrc/mjx_replay.cuh)."""
from __future__ import annotations

import json

import numpy as np

from .mjai_log import (ANKAN, CHI, DAHAI, DAIMINKAN, DORA, END_KYOKU, HORA, KAKAN, PON, REACH, REACH_ACCEPTED, RYUKYOKU,
                       START_KYOKU, TILE_NAMES, TSUMO)

START_GAME, END_GAME = 15, 16
TILE_ID = {name: i for i, name in enumerate(TILE_NAMES)}
_TYPES = {"start_kyoku": START_KYOKU, "tsumo": TSUMO, "dahai": DAHAI, "chi": CHI, "pon": PON, "daiminkan": DAIMINKAN,
          "kakan": KAKAN, "ankan": ANKAN, "dora": DORA, "reach": REACH, "reach_accepted": REACH_ACCEPTED, "hora": HORA,
          "ryukyoku": RYUKYOKU, "end_kyoku": END_KYOKU, "start_game": START_GAME, "end_game": END_GAME}


def _word(ty, actor=0, target=0, pai=37, tsumogiri=0, c=(0, 0, 0, 0)):
    return (ty | (actor & 3) << 8 | (target & 3) << 10 | (pai & 0xFF) << 12 | (tsumogiri & 1) << 20
            | c[0] << 24 | c[1] << 32 | c[2] << 40 | c[3] << 48)


def encode_events(events):
    """events: list of mjai dicts (a whole game, start_game .. end_game) -> (header words uint64 [n], kyoku payload uint64 [k, 9])"""
    hdr, pay = [], []
    for ev in events:
        ty = _TYPES[ev["type"]]
        t = TILE_ID
        if ty == START_KYOKU:
            kyoku_abs = (t[ev["bakaze"]] - 27) * 4 + ev["kyoku"] - 1
            hdr.append(_word(ty, pai=t[ev["dora_marker"]], c=(kyoku_abs, ev["honba"], ev["kyotaku"], ev["oya"])))
            sc = [int(x) & 0xFFFFFFFF for x in ev["scores"]]
            flat = bytes(t[x] for hand in ev["tehais"] for x in hand) + bytes(4)
            pay.append([sc[0] | sc[1] << 32, sc[2] | sc[3] << 32] + [int.from_bytes(flat[8 * k:8 * k + 8], "little") for k in range(7)])
        elif ty in (TSUMO,):
            hdr.append(_word(ty, ev["actor"], pai=t[ev["pai"]]))
        elif ty == DAHAI:
            hdr.append(_word(ty, ev["actor"], pai=t[ev["pai"]], tsumogiri=int(ev["tsumogiri"])))
        elif ty in (CHI, PON, DAIMINKAN):
            cons = [t[x] for x in ev["consumed"]] + [37] * (4 - len(ev["consumed"]))
            hdr.append(_word(ty, ev["actor"], ev["target"], t[ev["pai"]], c=cons))
        elif ty == KAKAN:
            cons = [t[x] for x in ev["consumed"]] + [37]
            hdr.append(_word(ty, ev["actor"], pai=t[ev["pai"]], c=cons))
        elif ty == ANKAN:
            hdr.append(_word(ty, ev["actor"], c=[t[x] for x in ev["consumed"]]))
        elif ty == DORA:
            hdr.append(_word(ty, pai=t[ev["dora_marker"]]))
        elif ty in (REACH, REACH_ACCEPTED):
            hdr.append(_word(ty, ev["actor"]))
        elif ty == HORA:
            hdr.append(_word(ty, ev["actor"], ev["target"]))
        else:  # ryukyoku, end_kyoku, start_game, end_game
            hdr.append(_word(ty))
    return np.array(hdr, dtype=np.uint64), np.array(pay, dtype=np.uint64).reshape(-1, 9)


def _swap_tile(name: str) -> str:
    """tile.rs:154-167 Tile::augment: manzu <-> pinzu, red fives stay red, souzu / honours / unknown unchanged"""
    if len(name) >= 2 and name[0].isdigit() and name[1] in "mp":
        return name[0] + ("p" if name[1] == "m" else "m") + name[2:]
    return name


def augment_events(events):
    """mjai/event.rs:187-217 Event::augment applied to a whole game: returns new event dicts, the input is not modified"""
    out = []
    for ev in events:
        e = dict(ev)
        for key in ("pai", "dora_marker", "bakaze"):
            if key in e:
                e[key] = _swap_tile(e[key])
        if "consumed" in e:
            e["consumed"] = [_swap_tile(x) for x in e["consumed"]]
        if "tehais" in e:
            e["tehais"] = [[_swap_tile(x) for x in hand] for hand in e["tehais"]]
        if e.get("ura_markers") is not None:
            e["ura_markers"] = [_swap_tile(x) for x in e["ura_markers"]]
        out.append(e)
    return out


def parse_log(text: str):
    return [json.loads(ln) for ln in text.splitlines() if ln.strip()]


def build_jobs(games, players_per_game):
    """games: list of event lists; players_per_game: list of player-id lists -> concatenated arrays for the replay kernels"""
    hdrs, pays, ev_off, ev_cnt, ky_off, players, job_game = [], [], [], [], [], [], []
    n_hdr = n_pay = 0
    for g, (events, pids) in enumerate(zip(games, players_per_game)):
        h, p = encode_events(events)
        for pid in pids:
            ev_off.append(n_hdr); ev_cnt.append(len(h)); ky_off.append(n_pay); players.append(pid); job_game.append(g)
        hdrs.append(h); pays.append(p)
        n_hdr += len(h); n_pay += len(p)
    cat = lambda xs, shape: np.concatenate(xs) if xs else np.zeros(shape, dtype=np.uint64)
    return dict(hdr=cat(hdrs, (0,)), kyoku=cat(pays, (0, 9)).reshape(-1), ev_off=np.array(ev_off, dtype=np.int32),
                ev_cnt=np.array(ev_cnt, dtype=np.int32), ky_off=np.array(ky_off, dtype=np.int32),
                players=np.array(players, dtype=np.uint8), job_game=np.array(job_game, dtype=np.int32))
