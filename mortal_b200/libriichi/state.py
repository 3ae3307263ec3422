"""libriichi.state — PlayerState / ActionCandidate on the CUDA environment.

Mirror of the PyO3 surface of state/player_state.rs:143-264, state/getter.rs:6-156, state/action.rs:11-89 and
state/obs_repr.rs:776-791: `PlayerState(player_id)`, `.update(mjai_json) -> ActionCandidate`, `.validate_reaction(mjai_json)`,
`.brief_info()`, `.encode_obs(version, at_kan_select) -> (ndarray f32 (C, 34), ndarray bool (46,))` and the getters. A PlayerState
is a device table record in single-seat mode (include/mjx.h mjx_state_*): every update is one launch of the very event handlers
self-play runs, the observation comes from the same encoder kernels (incl. the single-player tables of v4).

Beyond the PyO3 surface (the reference keeps them Rust-internal but asserts them in state/test.rs): `agari_points`,
`rule_based_agari`, `discard_candidates_aka`, `discard_candidates_with_unconditional_tenpai`, `clone`, `view`.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from .. import dataset_codec as DC
from ..mjai_log import (ANKAN, CHI, DAHAI, DAIMINKAN, HORA, KAKAN, PON, REACH, RYUKYOKU, TILE_NAMES)

CAN_FIELDS = ("can_discard", "can_chi_low", "can_chi_mid", "can_chi_high", "can_pon", "can_daiminkan", "can_kakan", "can_ankan",
              "can_riichi", "can_tsumo_agari", "can_ron_agari", "can_ryukyoku")  # action.rs:11-40, declaration order


class PlayerView(C.Structure):
    """include/mjx.h mjx_player_view"""
    _fields_ = [
        ("tehai", C.c_uint8 * 34), ("waits", C.c_uint8 * 34), ("dora_factor", C.c_uint8 * 34), ("tiles_seen", C.c_uint8 * 34),
        ("keep_shanten_discards", C.c_uint8 * 34), ("next_shanten_discards", C.c_uint8 * 34), ("forbidden_tiles", C.c_uint8 * 34),
        ("discarded_tiles", C.c_uint8 * 34),
        ("akas_seen", C.c_uint8 * 3), ("akas_in_hand", C.c_uint8 * 3),
        ("bakaze", C.c_uint8), ("jikaze", C.c_uint8), ("kyoku", C.c_uint8), ("honba", C.c_uint8), ("kyotaku", C.c_uint8),
        ("rank", C.c_uint8), ("oya", C.c_uint8), ("is_all_last", C.c_uint8),
        ("scores", C.c_int32 * 4),
        ("n_dora_indicators", C.c_uint8), ("dora_indicators", C.c_uint8 * 5),
        ("riichi_declared", C.c_uint8 * 4), ("riichi_accepted", C.c_uint8 * 4),
        ("at_turn", C.c_uint8), ("tiles_left", C.c_uint8),
        ("shanten", C.c_int8), ("real_time_shanten", C.c_int8),
        ("has_last_self_tsumo", C.c_uint8), ("last_self_tsumo", C.c_uint8), ("has_last_kawa_tile", C.c_uint8), ("last_kawa_tile", C.c_uint8),
        ("cans", C.c_uint32),
        ("n_ankan_candidates", C.c_uint8), ("ankan_candidates", C.c_uint8 * 3), ("n_kakan_candidates", C.c_uint8), ("kakan_candidates", C.c_uint8 * 3),
        ("chankan_chance", C.c_uint8), ("can_w_riichi", C.c_uint8), ("is_w_riichi", C.c_uint8), ("at_rinshan", C.c_uint8),
        ("at_ippatsu", C.c_uint8), ("at_furiten", C.c_uint8), ("to_mark_same_cycle_furiten", C.c_uint8), ("kans_on_board", C.c_uint8),
        ("is_menzen", C.c_uint8),
        ("n_chis", C.c_uint8), ("chis", C.c_uint8 * 4), ("n_pons", C.c_uint8), ("pons", C.c_uint8 * 4),
        ("n_minkans", C.c_uint8), ("minkans", C.c_uint8 * 4), ("n_ankans", C.c_uint8), ("ankans", C.c_uint8 * 4),
        ("doras_owned", C.c_uint8 * 4), ("doras_seen", C.c_uint8), ("tehai_len_div3", C.c_uint8), ("has_next_shanten_discard", C.c_uint8),
        ("kawa_len", C.c_uint8 * 4),
        ("viewer", C.c_uint8), ("pad_", C.c_uint8 * 3),
        ("err", C.c_int32),
    ]


class ActionCandidate:
    """state/action.rs:11-89"""

    def __init__(self, bits: int):
        self._bits = int(bits)
        for i, name in enumerate(CAN_FIELDS):
            setattr(self, name, bool((self._bits >> i) & 1))
        self.target_actor = (self._bits >> 16) & 0xFF

    can_chi = property(lambda s: s.can_chi_low or s.can_chi_mid or s.can_chi_high)
    can_kan = property(lambda s: s.can_daiminkan or s.can_kakan or s.can_ankan)
    can_agari = property(lambda s: s.can_tsumo_agari or s.can_ron_agari)
    can_pass = property(lambda s: s.can_chi or s.can_pon or s.can_daiminkan or s.can_ron_agari)
    can_act = property(lambda s: s.can_discard or s.can_chi or s.can_pon or s.can_kan or s.can_riichi or s.can_agari or s.can_ryukyoku)

    def __getitem__(self, key):  # dict-style access, as the oracle's test helper returns
        return getattr(self, key)

    def __repr__(self):
        on = [n for n in CAN_FIELDS if getattr(self, n)]
        return f"ActionCandidate({', '.join(on) or 'none'}, target_actor={self.target_actor})"


class _CudaBackend:
    """include/mjx.h mjx_state_* through libmjx.so (the product path; no CPU fallback)."""

    def __init__(self, device: int = 0):
        import torch

        from .. import _lib

        if not torch.cuda.is_available():
            raise _lib.MjxError("libriichi.state.PlayerState needs a CUDA device (mortal_b200 has no CPU fallback)")
        self.torch, self._lib, self.device = torch, _lib, device
        torch.cuda.set_device(device)
        _lib.init(device)
        self.L = _lib.load()

    def create(self, player_ids, version=4):
        ids = np.ascontiguousarray(player_ids, dtype=np.uint8)
        h = C.c_void_p()
        self._lib.check(self.L.mjx_state_create(C.byref(h), len(ids), ids.ctypes.data, version), "mjx_state_create")
        return h

    def destroy(self, h):
        self.L.mjx_env_destroy(h)

    def update(self, h, words, payload):
        cans = np.zeros(len(words), dtype=np.uint32)
        self._lib.check(self.L.mjx_state_update(h, words.ctypes.data, None if payload is None else payload.ctypes.data, cans.ctypes.data),
                        "mjx_state_update")
        return cans

    def view(self, h, index):
        v = PlayerView()
        self._lib.check(self.L.mjx_state_view(h, index, C.byref(v)), "mjx_state_view")
        return v

    def encode(self, h, n, version, kan):
        torch = self.torch
        rows = self.L.mjx_obs_rows(version)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        k = np.ascontiguousarray(kan, dtype=np.uint8)
        self._lib.check(self.L.mjx_env_set_obs_version(h, version), "mjx_env_set_obs_version")
        self._lib.check(self.L.mjx_state_rows(h, k.ctypes.data, st), "mjx_state_rows")
        cap = self.L.mjx_env_row_cap(h)
        obs = torch.empty((cap, rows, 34), dtype=torch.float32, device=f"cuda:{self.device}")
        self._lib.check(self.L.mjx_env_encode_obs(h, C.c_void_p(obs.data_ptr()), st), "mjx_env_encode_obs")
        from ..env import _CudaView

        masks = torch.as_tensor(_CudaView(self.L.mjx_env_masks(h), (cap, 46), "|u1"), device=obs.device)
        return obs[:n].cpu().numpy(), masks[:n].cpu().numpy().astype(bool)

    def query(self, h, index, what, args):
        a = np.zeros(8, dtype=np.int32)
        a[: len(args)] = args
        out = np.zeros(4, dtype=np.int32)
        self._lib.check(self.L.mjx_state_query(h, index, what, a.ctypes.data, out.ctypes.data), "mjx_state_query")
        return out

    def copy(self, dst, di, src, si):
        self._lib.check(self.L.mjx_state_copy(dst, di, src, si), "mjx_state_copy")


_backend = None


def set_backend(backend) -> None:
    """test hook: tests/emul_state.py injects the host-emulated backend; the default is the CUDA library"""
    global _backend
    _backend = backend


def get_backend():
    global _backend
    if _backend is None:
        _backend = _CudaBackend()
    return _backend


def _tile_name(t: int) -> str:
    return TILE_NAMES[t]


def reaction_from_word(w: int):
    """event word of a decoded reaction (csrc/mjx_step.cuh log_word) -> mjai dict, or {"type": "none"}"""
    ty = w & 0xFF
    actor, target, pai = (w >> 8) & 3, (w >> 10) & 3, (w >> 12) & 0xFF
    c = [(w >> s) & 0xFF for s in (24, 32, 40, 48)]
    t = TILE_NAMES
    if ty == DAHAI:
        return {"type": "dahai", "actor": actor, "pai": t[pai], "tsumogiri": bool((w >> 20) & 1)}
    if ty == REACH:
        return {"type": "reach", "actor": actor}
    if ty in (CHI, PON):
        return {"type": "chi" if ty == CHI else "pon", "actor": actor, "target": target, "pai": t[pai], "consumed": [t[c[0]], t[c[1]]]}
    if ty == DAIMINKAN:
        return {"type": "daiminkan", "actor": actor, "target": target, "pai": t[pai], "consumed": [t[c[0]], t[c[1]], t[c[2]]]}
    if ty == KAKAN:
        return {"type": "kakan", "actor": actor, "pai": t[pai], "consumed": [t[c[0]], t[c[1]], t[c[2]]]}
    if ty == ANKAN:
        return {"type": "ankan", "actor": actor, "consumed": [t[x] for x in c]}
    if ty == HORA:
        return {"type": "hora", "actor": actor, "target": target}
    if ty == RYUKYOKU:
        return {"type": "ryukyoku"}
    return {"type": "none"}


class PlayerState:
    def __init__(self, player_id: int):
        if not 0 <= int(player_id) <= 3:
            raise ValueError("player_id must be within 0..3")
        self._b = get_backend()
        self._pid = int(player_id)
        self._h = self._b.create([self._pid], 4)
        self._cans = ActionCandidate(self._pid << 16)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._b.destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- player_state.rs:157-171
    def update(self, mjai_json) -> ActionCandidate:
        ev = json.loads(mjai_json) if isinstance(mjai_json, str) else dict(mjai_json)
        ev.pop("meta", None)
        ev.pop("can_act", None)
        if ev.get("type") == "none":
            raise ValueError("cannot update the state with a `none` event")
        try:
            words, pay = DC.encode_events([ev])
        except KeyError as exc:
            raise ValueError(f"failed to parse event {ev}: {exc}") from None
        payload = np.ascontiguousarray(pay[0]) if len(pay) else None
        bits = self._b.update(self._h, np.ascontiguousarray(words[:1]), payload)[0]
        v = self.view()
        if v.err != 0:
            raise RuntimeError(f"on event {ev}: inconsistent state (mjx error code {v.err})")
        self._cans = ActionCandidate(bits)
        return self._cans

    def validate_reaction(self, mjai_json) -> None:
        """state/action.rs:93-227"""
        ev = json.loads(mjai_json) if isinstance(mjai_json, str) else dict(mjai_json)
        cans, v, ty = self._cans, self.view(), ev.get("type")
        tid = DC.TILE_ID

        def ensure(cond, msg):
            if not cond:
                raise ValueError(msg)

        def in_hand(tiles):  # action.rs:213-227
            for name in tiles:
                t = tid[name]
                d = (t - 34) * 9 + 4 if t >= 34 else t
                ensure(v.tehai[d] > 0, f"{name} is not in hand")
                if t >= 34:
                    ensure(v.akas_in_hand[t - 34], f"{name} is not in hand")

        last_kawa = _tile_name(v.last_kawa_tile) if v.has_last_kawa_tile else None
        if ty == "ryukyoku":
            ensure(cans.can_ryukyoku, "cannot ryukyoku")
            return
        if ty == "none":
            return
        ensure("actor" in ev, "action does not have actor and is not ryukyoku")
        ensure(ev["actor"] == self._pid, f"actor is {ev['actor']}, not self ({self._pid})")
        if ty == "dahai":
            ensure(cans.can_discard, "cannot discard")
            in_hand([ev["pai"]])
            if ev.get("tsumogiri"):
                ensure(v.has_last_self_tsumo, "tsumogiri but the player has not dealt any tile yet")
                ensure(_tile_name(v.last_self_tsumo) == ev["pai"], "cannot tsumogiri")
        elif ty == "reach":
            ensure(cans.can_riichi, "cannot riichi")
        elif ty == "chi":
            ensure((ev["target"] + 1) % 4 == ev["actor"], "chi from non-kamicha")
            ensure(last_kawa == ev["pai"], "chi target is not the last kawa tile")
            in_hand(ev["consumed"])
            de = lambda n: (tid[n] - 34) * 9 + 4 if tid[n] >= 34 else tid[n]
            a, b, t = de(ev["consumed"][0]), de(ev["consumed"][1]), de(ev["pai"])
            kind = "low" if t < min(a, b) else ("mid" if t < max(a, b) else "high")  # chi_type.rs:10-25
            ensure(getattr(cans, f"can_chi_{kind}"), f"cannot chi {kind}")
        elif ty in ("pon", "daiminkan"):
            ensure(ev["target"] != ev["actor"], f"{ty} from itself")
            ensure(last_kawa == ev["pai"], f"{ty} target is not the last kawa tile")
            ensure(cans.can_pon if ty == "pon" else cans.can_daiminkan, f"cannot {ty}")
            in_hand(ev["consumed"])
        elif ty == "kakan":
            ensure(cans.can_kakan, "cannot kakan")
            t = tid[ev["pai"]]
            d = (t - 34) * 9 + 4 if t >= 34 else t
            ensure(d in list(v.kakan_candidates)[: v.n_kakan_candidates], f"cannot kakan {ev['pai']}")
            in_hand([ev["pai"]])
        elif ty == "ankan":
            ensure(cans.can_ankan, "cannot ankan")
            t = tid[ev["consumed"][0]]
            d = (t - 34) * 9 + 4 if t >= 34 else t
            ensure(d in list(v.ankan_candidates)[: v.n_ankan_candidates], f"cannot ankan {TILE_NAMES[d]}")
            in_hand(ev["consumed"])
        elif ty == "hora":
            if ev.get("target") == self._pid:
                ensure(cans.can_tsumo_agari, "cannot tsumo agari")
            else:
                ensure(cans.can_ron_agari, "cannot ron agari")
        else:
            raise ValueError(f"unexpected action {ev}")

    def brief_info(self) -> str:
        """player_state.rs:173-264 (a human-readable digest; the single-player table is omitted)"""
        v = self.view()
        hand = " ".join(TILE_NAMES[t] for t in range(34) for _ in range(v.tehai[t]))
        waits = [TILE_NAMES[t] for t in range(34) if v.waits[t]]
        return (f"player (abs): {self._pid}\noya (rel): {v.oya}\nkyoku: {TILE_NAMES[v.bakaze]}{v.kyoku + 1}-{v.honba}\n"
                f"turn: {v.at_turn}\njikaze: {TILE_NAMES[v.jikaze]}\nscore (rel): {list(v.scores)}\ntehai: {hand}\n"
                f"tehai len: {v.tehai_len_div3}\nshanten: {v.shanten} (actual: {v.real_time_shanten})\nfuriten: {bool(v.at_furiten)}\n"
                f"waits: {waits}\ndora indicators: {[TILE_NAMES[v.dora_indicators[i]] for i in range(v.n_dora_indicators)]}\n"
                f"doras owned: {list(v.doras_owned)}\ndoras seen: {v.doras_seen}\naction candidates: {self._cans}\n"
                f"tiles left: {v.tiles_left}\n")

    # ---- obs_repr.rs:776-791
    def encode_obs(self, version: int, at_kan_select: bool):
        if version not in (1, 2, 3, 4):
            raise ValueError("unsupported version")
        obs, masks = self._b.encode(self._h, 1, version, [1 if at_kan_select else 0])
        return obs[0], masks[0]

    # ---- state/getter.rs:6-156
    def view(self) -> PlayerView:
        return self._b.view(self._h, 0)

    player_id = property(lambda s: s._pid)
    kyoku = property(lambda s: s.view().kyoku)
    honba = property(lambda s: s.view().honba)
    kyotaku = property(lambda s: s.view().kyotaku)
    is_oya = property(lambda s: s.view().oya == 0)
    tehai = property(lambda s: list(s.view().tehai))
    akas_in_hand = property(lambda s: [bool(x) for x in s.view().akas_in_hand])
    chis = property(lambda s: list(s.view().chis)[: s.view().n_chis])
    pons = property(lambda s: list(s.view().pons)[: s.view().n_pons])
    minkans = property(lambda s: list(s.view().minkans)[: s.view().n_minkans])
    ankans = property(lambda s: list(s.view().ankans)[: s.view().n_ankans])
    at_turn = property(lambda s: s.view().at_turn)
    shanten = property(lambda s: s.view().shanten)
    waits = property(lambda s: [bool(x) for x in s.view().waits])
    last_cans = property(lambda s: s._cans)
    can_w_riichi = property(lambda s: bool(s.view().can_w_riichi))
    self_riichi_declared = property(lambda s: bool(s.view().riichi_declared[0]))
    self_riichi_accepted = property(lambda s: bool(s.view().riichi_accepted[0]))
    at_furiten = property(lambda s: bool(s.view().at_furiten))

    def last_self_tsumo(self):
        v = self.view()
        return TILE_NAMES[v.last_self_tsumo] if v.has_last_self_tsumo else None

    def last_kawa_tile(self):
        v = self.view()
        return TILE_NAMES[v.last_kawa_tile] if v.has_last_kawa_tile else None

    def ankan_candidates(self):
        v = self.view()
        return [TILE_NAMES[v.ankan_candidates[i]] for i in range(v.n_ankan_candidates)]

    def kakan_candidates(self):
        v = self.view()
        return [TILE_NAMES[v.kakan_candidates[i]] for i in range(v.n_kakan_candidates)]

    # ---- Rust-internal API that state/test.rs asserts
    def agari_points(self, is_ron: bool, ura=()):
        u = [DC.TILE_ID[t] if isinstance(t, str) else int(t) for t in ura]
        out = self._b.query(self._h, 0, 0, [int(is_ron), len(u)] + u)
        if not out[3]:
            raise RuntimeError("agari_points: not an agari hand")
        return dict(ron=int(out[0]), tsumo_ko=int(out[1]), tsumo_oya=int(out[2]))

    def rule_based_agari(self) -> bool:
        return bool(self._b.query(self._h, 0, 1, [])[0])

    def discard_candidates(self, unconditional_tenpai: bool = False) -> np.ndarray:
        out = self._b.query(self._h, 0, 3 if unconditional_tenpai else 2, [])
        m = (int(out[0]) & 0xFFFFFFFF) | ((int(out[1]) & 0xFFFFFFFF) << 32)
        return np.array([(m >> i) & 1 for i in range(37)], dtype=bool)

    def decode_action(self, action: int, kan_select_action: int = -1):
        """agent/mortal.rs:338-573: action id -> the reaction event (mjai dict) this seat would send"""
        out = self._b.query(self._h, 0, 4, [int(action), int(kan_select_action)])
        if out[2] != 0:
            raise ValueError(f"action {action} cannot be decoded in this state (mjx error code {int(out[2])})")
        w = (int(out[0]) & 0xFFFFFFFF) | ((int(out[1]) & 0xFFFFFFFF) << 32)
        return reaction_from_word(w)

    def clone(self) -> "PlayerState":
        other = PlayerState(self._pid)
        self._b.copy(other._h, 0, self._h, 0)
        other._cans = ActionCandidate(self._cans._bits)
        return other

    @classmethod
    def from_log(cls, player_id: int, lines) -> "PlayerState":
        ps = cls(player_id)
        for ln in lines:
            ps.update(ln)
        return ps
