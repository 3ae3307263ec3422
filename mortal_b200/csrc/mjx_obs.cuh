// mortal_b200 — observation encoder: table record + seat -> (C, 34) f32 planes.
//
// Contract: libriichi state/obs_repr.rs:126-630 (row map for version 4 in SURVEY.md §8 a17),
// helpers obs_repr.rs:694-774, discard_candidates_with_unconditional_tenpai agent_helper.rs:100-197.
// The reference fills a heap array through a running row cursor, one PlayerState per seat. Here the
// whole (1012, 34) tile is assembled in shared memory from the table record (public part shared by
// the four perspectives, rotated on the fly) by the warps of one CTA, each warp owning a set of
// feature sections with lanes mapped to the 34 tile columns, and then leaves the SM as ONE bulk
// asynchronous copy (TMA, cp.async.bulk shared->global) of 137,632 contiguous bytes.
#pragma once
#include "mjx_step.cuh"

namespace mjx {

constexpr int OBS_ROWS_V4 = 1012;
constexpr int OBS_COLS = 34;

struct EncCtx {
    const TableState* S;
    Tables T;
    float* tile;    // [rows * 34], zero-filled by the caller
    int seat;       // perspective (absolute seat)
    bool kan_select;
    int lane, warp, nwarps;
    const u8* dora_factor;  // [34]
    int row_lo, row_hi;     // this CTA builds obs rows [row_lo, row_hi); tile points at row_lo
};

// An observation is built as two half-tiles (rows [0,522) and [522,1012)) by different CTAs so that three
// CTAs fit one SM and tile assembly overlaps the bulk stores of the others. The split falls on a section
// boundary: sections 0-5 (+2a) live in the first half, 6-10 (+2b) in the second.
constexpr int OBS_SPLIT_ROW = 522;

// rows are addressed absolutely; ENC_AT maps them into the CTA's window
#define ENC_AT(e, row, col) (e).tile[((row) - (e).row_lo) * 34 + (col)]
#define ENC_HALF(e, first_half) ((first_half) ? (e).row_lo < OBS_SPLIT_ROW : (e).row_hi > OBS_SPLIT_ROW)
#ifdef MJX_HOST_EMUL
#define ENC_SECTION(e, k, first_half) ENC_HALF(e, first_half)
#define ENC_FILL(e, row, v) do { for (int c_ = 0; c_ < 34; c_++) ENC_AT(e, row, c_) = (v); } while (0)
#define ENC_ASSIGN(e, row, col, v) do { ENC_AT(e, row, col) = (v); } while (0)
#define ENC_SYNCWARP() ((void)0)
#else
#define ENC_SECTION(e, k, first_half) (ENC_HALF(e, first_half) && (((k) % (e).nwarps) == (e).warp))
#define ENC_FILL(e, row, v) do { ENC_AT(e, row, (e).lane) = (v); if ((e).lane < 2) ENC_AT(e, row, 32 + (e).lane) = (v); } while (0)
#define ENC_ASSIGN(e, row, col, v) do { if ((e).lane == 0) ENC_AT(e, row, col) = (v); } while (0)
#define ENC_SYNCWARP() __syncwarp()
#endif

MJX_D int rel_to_abs(int seat, int rel) { return (seat + rel) & 3; }

// a seat's pond as the perspective `p` sees it: optional start-of-kyoku pad (update.rs:819-824) + items
struct KawaView {
    const KawaItem* items;
    int n_items;
    int pad;  // 0 or 1
    MJX_DM int len() const { return n_items + pad; }
    // returns nullptr for a None slot
    MJX_DM const KawaItem* at(int i) const {
        if (i < pad) return nullptr;
        const KawaItem* k = items + (i - pad);
        return k->tile == T_NONE ? nullptr : k;
    }
};

MJX_D KawaView kawa_view(const TableState* S, int p, int abs_seat) {
    KawaView v;
    v.items = S->pub[abs_seat].kawa;
    v.n_items = S->pub[abs_seat].kawa_len;
    int rel_s = (abs_seat - p) & 3, rel_oya = (S->oya - p) & 3;
    v.pad = rel_s < rel_oya ? 1 : 0;
    return v;
}

// obs_repr.rs:694-712 over an explicit tile list accessor
template <typename F>
MJX_D void enc_tile_set(EncCtx& e, int row, int n, F get) {
    // executed by one lane: per-tile running counts
    if (e.lane == 0) {
        u8 counts[34];
        for (int i = 0; i < 34; i++) counts[i] = 0;
        for (int i = 0; i < n; i++) {
            int tile = get(i);
            int tid = deaka(tile);
            ENC_AT(e, row + counts[tid], tid) = 1.f;
            counts[tid]++;
        }
    }
    ENC_SYNCWARP();
    for (int i = 0; i < n; i++) {
        int tile = get(i);
        if (is_aka(tile)) ENC_FILL(e, row + 4 + (tile - T_5MR), 1.f);
    }
}

// agent_helper.rs:100-197 as a 34-bit mask (aka folded back as discard_candidates_with_unconditional_tenpai does)
MJX_DN u64 unconditional_tenpai_discards(EncCtx& e, const Ctx& c) {
    const TableState* S = e.S;
    const int p = e.seat;
    const SeatPrivate& P = S->priv[p];
    const bool has_next = (P.flags & PF_HAS_NEXT_SHANTEN_DISCARD) != 0;
    if (S->tiles_left == 0 || P.shanten > 1 || (P.shanten == 1 && !has_next)) return 0;
    const bool racc = (S->riichi_accepted >> p) & 1;
    if (P.last_self_tsumo != T_NONE) {
        if ((P.waits >> deaka(P.last_self_tsumo)) & 1) return 0;
        if (racc) return (P.flags & PF_AT_FURITEN) ? 0 : (1ull << deaka(P.last_self_tsumo));
    } else if (shanten_all(e.T, P.tehai, P.tehai_len_div3) == -1) {
        return 0;
    }
    const u64 cand = (P.shanten == 1 ? P.next_shanten : P.keep_shanten) & ~P.forbidden;
    u64 result = 0;
    const int len = P.tehai_len_div3;
    const HandSig base = hand_sig(P.tehai);
    for (u64 rest = cand; rest; rest &= rest - 1) {
        const int discard = mjx_ffsll(rest) - 1;
        const HandSig b1 = sig_variant(base, discard, -1, P.tehai[discard]);
        // every tsumo tile evaluated by its own lane: code bit0 = completes the hand, bit1 = has yaku
        u64 wins, yaku;
        tile_eval2(c, true, [&](int t) {
            int n = P.tehai[t] - (t == discard ? 1 : 0);
            if (t == discard || n == 4) return 0;
            if (shanten_all_sig(e.T, sig_variant(b1, t, +1, n), len) > -1) return 0;
            int code = 1;
            if (!((P.discarded >> t) & 1) && (S->public_seen[t] + P.tehai[t]) < 4) {
                u8 th[34];
                for (int i = 0; i < 34; i++) th[i] = P.tehai[i];
                th[discard] -= 1;
                th[t] += 1;
                if (has_yaku(e.T, make_query(S, p, th, t, true))) code |= 2;
            }
            return code;
        }, wins, yaku);
        // reference loop order: ascending tsumo; a furiten winning tile voids the discard, otherwise
        // the discard qualifies as soon as one live winning tile has a yaku
        bool ok = false;
        if (!(wins & P.discarded)) ok = yaku != 0;
        else {
            // furiten tile found at position f: result is false regardless (ret[discard] = false; break)
            ok = false;
        }
        if (ok) result |= 1ull << discard;
    }
    return result;
}

// Encodes version-4 rows 0..888 (+ leaves 889..1011, the single-player block, to encode_sp).
// All warps of the CTA call this with the same arguments; the tile must be zero on entry.
MJX_DN void encode_obs_v4(EncCtx& e, const Ctx& c, u64* mask_out) {
    const TableState* S = e.S;
    const int p = e.seat;
    const SeatPrivate& P = S->priv[p];
    const u16 cans = P.cans;
    const u8* df = e.dora_factor;

    // ---- section 0: hand (rows 0-6)
    if (ENC_SECTION(e, 0, true)) {
        MJX_FOR_TILES(e, t) {
            for (int n = 0; n < P.tehai[t]; n++) ENC_AT(e, n, t) = 1.f;
        }
        for (int k = 0; k < 3; k++) if ((P.akas_in_hand >> k) & 1) ENC_FILL(e, 4 + k, 1.f);
    }
    // ---- section 1: scalars (rows 7-27)
    if (ENC_SECTION(e, 1, true)) {
        int rank = 0;
        for (int i = 0; i < 4; i++) {
            i32 sc = S->scores[rel_to_abs(p, i)];
            ENC_FILL(e, 7 + 2 * i, (float)min(max(sc, 0), 100000) / 100000.f);
            ENC_FILL(e, 8 + 2 * i, (float)min(max(sc, 0), 30000) / 30000.f);
        }
        for (int s = 0; s < 4; s++)  // rankings.rs:8-22: stable by seat
            if (s != p && (S->scores[s] > S->scores[p] || (S->scores[s] == S->scores[p] && s < p))) rank++;
        ENC_FILL(e, 15 + rank, 1.f);
        const int kyoku_in_wind = S->kyoku & 3, bakaze = T_E + S->kyoku / 4;
        ENC_FILL(e, 19 + kyoku_in_wind, 1.f);
        ENC_FILL(e, 23, (float)min((int)S->honba, 10) / 10.f);
        ENC_FILL(e, 24, (float)min((int)S->kyotaku, 10) / 10.f);
        ENC_ASSIGN(e, 25, bakaze, 1.f);
        ENC_ASSIGN(e, 26, T_E + ((p + 4 - S->oya) & 3), 1.f);
        int gk = min(bakaze - T_E, 1) * 4 + kyoku_in_wind;
        ENC_FILL(e, 27, (float)min(gk, 7) / 7.f);
    }
    // ---- section 2a: dora indicators (rows 28-34)
    if (ENC_SECTION(e, 2, true)) enc_tile_set(e, 28, S->n_dora, [&](int i) { return dora_indicator(S, i); });
    // ---- section 2b: counters (rows 717-722)
    if (ENC_SECTION(e, 5, false)) {
        ENC_FILL(e, 717, (float)S->tiles_left / 69.f);
        int seen_doras = mjx_popc((u32)(S->akas_public | P.akas_in_hand));
        for (int t = 0; t < 34; t++) seen_doras += (S->public_seen[t] + P.tehai[t]) * df[t];
        for (int i = 0; i < 4; i++) {
            const int s = rel_to_abs(p, i);
            const SeatPublic& U = S->pub[s];
            int n = 0;
            if (i == 0) {
                n = mjx_popc((u32)P.akas_in_hand);
                for (int t = 0; t < 34; t++) n += P.tehai[t] * df[t];
            }
            for (int f = 0; f < U.n_fuuro; f++)
                for (int j = 0; j < 4; j++) {
                    int t = U.fuuro[f][j];
                    if (t != T_NONE) n += df[deaka(t)] + (is_aka(t) ? 1 : 0);
                }
            for (int j = 0; j < U.n_ankan; j++) {
                int t = U.ankan[j];
                n += 4 * df[t] + ((t == T_5M || t == T_5P || t == T_5S) ? 1 : 0);
            }
            ENC_FILL(e, 718 + i, (float)min(n, 12) / 12.f);
        }
        int unseen = (S->n_dora * 4 + 3 - seen_doras) & 0xFF;
        ENC_FILL(e, 722, (float)min(unseen, 23) / 23.f);
    }
    // max kawa length over the four ponds as this seat sees them (obs_repr.rs:221)
    int max_kawa_len = 0;
    for (int s = 0; s < 4; s++) max_kawa_len = max(max_kawa_len, kawa_view(S, p, s).len());

    // ---- section 3: own pond (rows 35-131)
    if (ENC_SECTION(e, 3, true)) {
        const KawaView kv = kawa_view(S, p, p);
        const int len = kv.len();
        for (int pass = 0; pass < 2; pass++) {
            const int slots = pass == 0 ? 6 : 18, base = pass == 0 ? 35 : 59;
            for (int j = 0; j < min(len, slots); j++) {
                const KawaItem* k = kv.at(pass == 0 ? j : len - 1 - j);
                if (!k) continue;
                const int row = base + 4 * j;
                for (int q = 0; q < 4; q++) if (k->kan[q] != T_NONE) ENC_ASSIGN(e, row, k->kan[q], 1.f);
                ENC_ASSIGN(e, row + 1, deaka(k->tile), 1.f);
                if (is_aka(k->tile)) ENC_FILL(e, row + 2, 1.f);
                if (k->flags & SF_DORA) ENC_FILL(e, row + 3, 1.f);
            }
        }
        for (int turn = 0; turn < len; turn++) {
            const KawaItem* k = kv.at(turn);
            if (k) ENC_ASSIGN(e, 131, deaka(k->tile), expf(-0.2f * (float)(max_kawa_len - 1 - turn)));
        }
    }
    // ---- sections 4-6: the three opponents' ponds (rows 132-716)
    for (int rel = 1; rel < 4; rel++) {
        if (!ENC_SECTION(e, rel == 3 ? 0 : 3 + rel, rel != 3)) continue;
        const KawaView kv = kawa_view(S, p, rel_to_abs(p, rel));
        const int len = kv.len();
        const int sec = 132 + 195 * (rel - 1);
        for (int pass = 0; pass < 2; pass++) {
            const int slots = pass == 0 ? 6 : 18, base = sec + (pass == 0 ? 0 : 48);
            for (int j = 0; j < min(len, slots); j++) {
                const KawaItem* k = kv.at(pass == 0 ? j : len - 1 - j);
                if (!k) continue;
                const int row = base + 8 * j;
                if (k->flags & SF_HAS_CHIPON) {
                    ENC_ASSIGN(e, row, min(k->consumed[0], k->consumed[1]), 1.f);
                    ENC_ASSIGN(e, row + 1, max(k->consumed[0], k->consumed[1]), 1.f);
                }
                for (int q = 0; q < 4; q++) if (k->kan[q] != T_NONE) ENC_ASSIGN(e, row + 2, k->kan[q], 1.f);
                ENC_ASSIGN(e, row + 3, deaka(k->tile), 1.f);
                if (is_aka(k->tile)) ENC_FILL(e, row + 4, 1.f);
                if (k->flags & SF_DORA) ENC_FILL(e, row + 5, 1.f);
                if (k->flags & SF_TEDASHI) ENC_FILL(e, row + 6, 1.f);
                if (k->flags & SF_RIICHI) ENC_FILL(e, row + 7, 1.f);
            }
        }
        for (int turn = 0; turn < len; turn++) {
            const KawaItem* k = kv.at(turn);
            if (!k) continue;
            const float v = expf(-0.2f * (float)(max_kawa_len - 1 - turn));
            const int tid = deaka(k->tile);
            ENC_ASSIGN(e, sec + 192, tid, v);
            if (k->flags & SF_TEDASHI) ENC_ASSIGN(e, sec + 193, tid, v);
            if (k->flags & SF_RIICHI) ENC_ASSIGN(e, sec + 194, tid, v);
        }
    }
    // ---- section 7: kawa overview (rows 723-750)
    if (ENC_SECTION(e, 1, false)) {
        for (int i = 0; i < 4; i++) {
            const SeatPublic& U = S->pub[rel_to_abs(p, i)];
            // real discards only, in order (update.rs:336)
            int idx[KAWA_CAP], n = 0;
            for (int j = 0; j < U.kawa_len; j++) if (U.kawa[j].tile != T_NONE) idx[n++] = j;
            enc_tile_set(e, 723 + 7 * i, n, [&](int j) { return (int)U.kawa[idx[j]].tile; });
        }
    }
    // ---- section 8: melds (rows 751-834)
    if (ENC_SECTION(e, 2, false)) {
        for (int i = 0; i < 4; i++) {
            const SeatPublic& U = S->pub[rel_to_abs(p, i)];
            for (int f = 0; f < U.n_fuuro; f++) {
                const int row = 751 + 20 * i + 5 * f;
                for (int j = 0; j < 4; j++) {
                    int t = U.fuuro[f][j];
                    if (t == T_NONE) continue;
                    int tid = deaka(t), dup = 0;
                    for (int q = 0; q < j; q++) dup += U.fuuro[f][q] != T_NONE && deaka(U.fuuro[f][q]) == tid;
                    ENC_ASSIGN(e, row + dup, tid, 1.f);  // first still-zero plane (obs_repr.rs:305-308)
                    if (is_aka(t)) ENC_FILL(e, row + 4, 1.f);
                }
            }
            for (int j = 0; j < U.n_ankan; j++) ENC_ASSIGN(e, 831 + i, U.ankan[j], 1.f);
        }
    }
    // ---- section 9: seen tiles, key discards, riichi / wait status (rows 835-873)
    if (ENC_SECTION(e, 3, false)) {
        MJX_FOR_TILES(e, t) {
            ENC_AT(e, 835, t) = (float)(S->public_seen[t] + P.tehai[t]) / 4.f;
            if ((P.waits >> t) & 1) ENC_AT(e, 860, t) = 1.f;
        }
        for (int rel = 1; rel < 4; rel++) {
            const SeatPublic& U = S->pub[rel_to_abs(p, rel)];
            if (U.last_tedashi_flags & SF_VALID) {
                const int row = 836 + 3 * (rel - 1);
                ENC_ASSIGN(e, row, deaka(U.last_tedashi_tile), 1.f);
                if (is_aka(U.last_tedashi_tile)) ENC_FILL(e, row + 1, 1.f);
                if (U.last_tedashi_flags & SF_DORA) ENC_FILL(e, row + 2, 1.f);
            }
            if (U.riichi_flags & SF_VALID) {
                const int row = 845 + 3 * (rel - 1);
                ENC_ASSIGN(e, row, deaka(U.riichi_tile), 1.f);
                if (is_aka(U.riichi_tile)) ENC_FILL(e, row + 1, 1.f);
                if (U.riichi_flags & SF_DORA) ENC_FILL(e, row + 2, 1.f);
            }
            if ((S->riichi_declared >> rel_to_abs(p, rel)) & 1) ENC_FILL(e, 854 + rel - 1, 1.f);
            if ((S->riichi_accepted >> rel_to_abs(p, rel)) & 1) ENC_FILL(e, 857 + rel - 1, 1.f);
        }
        if (P.flags & PF_AT_FURITEN) ENC_FILL(e, 861, 1.f);
        ENC_FILL(e, 862 + min(max((int)P.shanten, 0), 6), 1.f);
        if ((S->riichi_accepted >> p) & 1) ENC_FILL(e, 869, 1.f);
        if (e.kan_select) ENC_FILL(e, 870, 1.f);
        if (cans & CAN_PASS) {
            const int tile = S->last_kawa_tile, tid = deaka(tile);
            ENC_ASSIGN(e, 871, tid, 1.f);
            if (is_aka(tile)) ENC_FILL(e, 872, 1.f);
            if (df[tid] > 0) ENC_FILL(e, 873, 1.f);
        }
    }
    // ---- section 10: the action block (rows 874-888) + legal mask
    const bool second = ENC_HALF(e, false);
    const u64 discards = (second && (cans & CAN_DISCARD)) ? discard_candidates(c, p) : 0;  // collective, every warp
    if (mask_out && second) *mask_out = legal_mask(c, p, e.kan_select, discards);
    if (ENC_SECTION(e, 4, false)) {
        if (cans & CAN_DISCARD) {
            u64 d34 = (discards & ((1ull << 34) - 1)) | (((discards >> 34) & 1) << 4) | (((discards >> 35) & 1) << 13) |
                      (((discards >> 36) & 1) << 22);
            u64 ut = 0;
            if (P.shanten <= 1) ut = unconditional_tenpai_discards(e, c);
            MJX_FOR_TILES(e, t) {
                if ((d34 >> t) & 1) ENC_AT(e, 874, t) = 1.f;
                if ((P.keep_shanten >> t) & 1) ENC_AT(e, 875, t) = 1.f;
                if ((P.next_shanten >> t) & 1) ENC_AT(e, 876, t) = 1.f;
                if ((ut >> t) & 1) ENC_AT(e, 877, t) = 1.f;
            }
            if ((S->riichi_declared >> p) & 1) ENC_FILL(e, 878, 1.f);
        }
        if (cans & CAN_RIICHI) ENC_FILL(e, 879, 1.f);
        if (cans & CAN_CHI_LOW) ENC_FILL(e, 880, 1.f);
        if (cans & CAN_CHI_MID) ENC_FILL(e, 881, 1.f);
        if (cans & CAN_CHI_HIGH) ENC_FILL(e, 882, 1.f);
        if (cans & CAN_PON) ENC_FILL(e, 883, 1.f);
        if (cans & CAN_DAIMINKAN) ENC_FILL(e, 884, 1.f);
        MJX_FOR_TILES(e, t) {
            if ((cans & CAN_ANKAN) && ((P.ankan_cand >> t) & 1)) ENC_AT(e, 885, t) = 1.f;
            if ((cans & CAN_KAKAN) && ((P.kakan_cand >> t) & 1)) ENC_AT(e, 886, t) = 1.f;
        }
        if (cans & CAN_AGARI) ENC_FILL(e, 887, 1.f);
        if (cans & CAN_RYUKYOKU) ENC_FILL(e, 888, 1.f);
    }
}

}  // namespace mjx
