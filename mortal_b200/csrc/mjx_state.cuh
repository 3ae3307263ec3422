// mortal_b200 — libriichi.state.PlayerState on device (state/player_state.rs:143-264, state/getter.rs, state/update.rs).
//
// A PlayerState is a table record in single-seat mode (TableState.viewer1 = seat + 1): the public part is maintained exactly as
// in self-play, of the private parts only the viewer's is known; events arrive as the 64-bit words of the event log
// (csrc/mjx_step.cuh log_word, csrc/mjx_replay.cuh apply_event), other seats' hidden tiles as `?`. One warp per state.
#pragma once
#include "../../include/mjx.h"
#include "mjx_sp.cuh"
#include "mjx_replay.cuh"

namespace mjx {

// state/getter.rs + the fields state/test.rs asserts: everything a caller may read back, derived like the encoder derives it
MJX_DN void state_view(const Ctx& c, int p, mjx_player_view* v) {
    const TableState* S = c.S;
    const SeatPrivate& P = S->priv[p];
    for (int t = 0; t < 34; t++) {
        v->tehai[t] = P.tehai[t];
        v->waits[t] = (u8)((P.waits >> t) & 1);
        v->dora_factor[t] = c.df[t];
        v->tiles_seen[t] = (u8)(S->public_seen[t] + P.tehai[t]);
        v->keep_shanten_discards[t] = (u8)((P.keep_shanten >> t) & 1);
        v->next_shanten_discards[t] = (u8)((P.next_shanten >> t) & 1);
        v->forbidden_tiles[t] = (u8)((P.forbidden >> t) & 1);
        v->discarded_tiles[t] = (u8)((P.discarded >> t) & 1);
    }
    for (int i = 0; i < 3; i++) {
        v->akas_in_hand[i] = (u8)((P.akas_in_hand >> i) & 1);
        v->akas_seen[i] = (u8)(((S->akas_public | P.akas_in_hand) >> i) & 1);
    }
    v->bakaze = (u8)(T_E + S->kyoku / 4);
    v->jikaze = (u8)(T_E + ((p + 4 - S->oya) & 3));
    v->kyoku = (u8)(S->kyoku % 4);
    v->honba = S->honba; v->kyotaku = S->kyotaku;
    int rank = 0;
    for (int s = 0; s < 4; s++)  // rankings.rs:8-22: stable by seat
        if (s != p && (S->scores[s] > S->scores[p] || (S->scores[s] == S->scores[p] && s < p))) rank++;
    v->rank = (u8)rank;
    v->oya = (u8)((S->oya + 4 - p) & 3);
    v->is_all_last = (u8)(S->kyoku / 4 == 0 ? 0 : (S->kyoku / 4 == 1 ? (S->kyoku % 4 == 3) : 1));
    for (int i = 0; i < 4; i++) {
        const int s = (p + i) & 3;
        v->scores[i] = S->scores[s];
        v->riichi_declared[i] = (u8)((S->riichi_declared >> s) & 1);
        v->riichi_accepted[i] = (u8)((S->riichi_accepted >> s) & 1);
        v->kawa_len[i] = S->pub[s].kawa_len;
    }
    v->n_dora_indicators = S->n_dora;
    for (int k = 0; k < 5; k++) v->dora_indicators[k] = k < S->n_dora ? S->wall[60 - k] : (u8)T_NONE;
    v->at_turn = P.at_turn; v->tiles_left = S->tiles_left;
    v->shanten = P.shanten;
    v->real_time_shanten = (i8)real_time_shanten(c.T, S, p);
    v->has_last_self_tsumo = P.last_self_tsumo != T_NONE; v->last_self_tsumo = P.last_self_tsumo;
    v->has_last_kawa_tile = S->last_kawa_tile != T_NONE; v->last_kawa_tile = S->last_kawa_tile;
    v->cans = (u32)P.cans | ((u32)P.target_actor << 16);
    v->n_ankan_candidates = v->n_kakan_candidates = 0;
    for (int t = 0; t < 34; t++) {
        if (((P.ankan_cand >> t) & 1) && v->n_ankan_candidates < 3) v->ankan_candidates[v->n_ankan_candidates++] = (u8)t;
        if (((P.kakan_cand >> t) & 1) && v->n_kakan_candidates < 3) v->kakan_candidates[v->n_kakan_candidates++] = (u8)t;
    }
    v->chankan_chance = (P.flags & PF_CHANKAN_CHANCE) != 0; v->can_w_riichi = (P.flags & PF_CAN_W_RIICHI) != 0;
    v->is_w_riichi = (P.flags & PF_IS_W_RIICHI) != 0; v->at_rinshan = (P.flags & PF_AT_RINSHAN) != 0;
    v->at_ippatsu = (P.flags & PF_AT_IPPATSU) != 0; v->at_furiten = (P.flags & PF_AT_FURITEN) != 0;
    v->to_mark_same_cycle_furiten = (P.flags & PF_MARK_SAME_CYCLE_FURITEN) != 0;
    v->kans_on_board = S->kans; v->is_menzen = (P.flags & PF_IS_MENZEN) != 0;
    v->n_chis = P.n_chis; v->n_pons = P.n_pons; v->n_minkans = P.n_minkans; v->n_ankans = P.n_ankans;
    for (int i = 0; i < 4; i++) { v->chis[i] = P.chis[i]; v->pons[i] = P.pons[i]; v->minkans[i] = P.minkans[i]; v->ankans[i] = P.ankans[i]; }
    // doras: recount (the reference's own invariant, state/test.rs:30-58)
    int seen_doras = 0, own_doras = mjx_popc((u32)P.akas_in_hand);
    for (int i = 0; i < 3; i++) seen_doras += ((S->akas_public | P.akas_in_hand) >> i) & 1;
    for (int k = 0; k < S->n_dora; k++) {
        const int d = tile_next(S->wall[60 - k]);
        seen_doras += S->public_seen[d] + P.tehai[d];
        own_doras += P.tehai[d];
    }
    for (int i = 0; i < 4; i++) {
        const SeatPublic& U = S->pub[(p + i) & 3];
        int n = i == 0 ? own_doras : 0;
        for (int f = 0; f < U.n_fuuro; f++)
            for (int j = 0; j < 4; j++) { const int t = U.fuuro[f][j]; if (t != T_NONE) n += c.df[deaka(t)] + (is_aka(t) ? 1 : 0); }
        for (int j = 0; j < U.n_ankan; j++) { const int t = U.ankan[j]; n += 4 * c.df[t] + ((t == T_5M || t == T_5P || t == T_5S) ? 1 : 0); }
        v->doras_owned[i] = (u8)n;
    }
    v->doras_seen = (u8)seen_doras;
    v->tehai_len_div3 = P.tehai_len_div3;
    v->has_next_shanten_discard = (P.flags & PF_HAS_NEXT_SHANTEN_DISCARD) != 0;
    v->viewer = (u8)p; v->err = S->err;
}

}  // namespace mjx
