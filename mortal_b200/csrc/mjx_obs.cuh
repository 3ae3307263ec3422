// mortal_b200 — observation encoder: table record + seat -> (C, 34) f32 planes.
//
// Contract: libriichi state/obs_repr.rs:126-630 (row map for version 4 in SURVEY.md §8 a17),
// helpers obs_repr.rs:694-774, discard_candidates_with_unconditional_tenpai agent_helper.rs:100-197.
// The reference fills a heap array through a running row cursor, one PlayerState per seat. An observation
// is ~97% zeros and almost every non-zero is exactly 1, so here a warp first derives a COMPACT form of one
// observation from its staged copy of the table record (public part shared by the four perspectives,
// rotated on the fly):
//   * `bm[row]`  — one 34-bit column mask per row for the cells that are 1.0 (rows 0..888), and
//   * `sv[slot]` — 34 floats for each of the 28 rows that can hold other values (scores, counters, exp-decay
//                  planes, seen/4), a static row -> slot map,
// (k_encode_features, ~11 KB per observation, L2-resident), and a second, purely streaming kernel
// (k_encode_store) materialises it slice by slice (OBS_SLICE_ROWS rows) in double-buffered shared-memory tiles
// that leave the SM as bulk asynchronous copies (TMA, cp.async.bulk shared->global). Warps are independent
// pipelines; there is no CTA-wide barrier anywhere.
#pragma once
#include "mjx_step.cuh"

namespace mjx {

constexpr int OBS_ROWS_V4 = 1012;
constexpr int SP_ROW0 = 889;  // first row of the single-player block (obs_repr.rs:561)
constexpr int OBS_COLS = 34;
// rows per slice: even, so that every slice starts on a 16-byte boundary (2 rows = 272 B) as bulk copies require
constexpr int OBS_SLICE_ROWS = 46;
constexpr int OBS_N_SLICES = (OBS_ROWS_V4 + OBS_SLICE_ROWS - 1) / OBS_SLICE_ROWS;  // 22, exact
static_assert(OBS_SLICE_ROWS % 2 == 0, "slice boundaries must be 16-byte aligned");
constexpr int OBS_BM_ROWS = 892;          // mask rows kept per observation (>= SP_ROW0, multiple of 4)
constexpr int OBS_N_SPECIAL = 28;         // rows that can hold values other than 0/1
constexpr u64 OBS_FULL_ROW = (1ull << 34) - 1;

// static map of the value rows (v4): scores 7-14, honba 23, kyotaku 24, game progress 27, own pond decay 131,
// opponents' pond decay 324-326 / 519-521 / 714-716, counters 717-722, seen/4 835
MJX_HD int enc_special_slot(int row) {
    if (row >= 7 && row <= 14) return row - 7;
    if (row == 23) return 8;
    if (row == 24) return 9;
    if (row == 27) return 10;
    if (row == 131) return 11;
    if (row >= 324 && row <= 326) return 12 + (row - 324);
    if (row >= 519 && row <= 521) return 15 + (row - 519);
    if (row >= 714 && row <= 716) return 18 + (row - 714);
    if (row >= 717 && row <= 722) return 21 + (row - 717);
    if (row == 835) return 27;
    return -1;
}
MJX_HD int enc_special_row(int slot) {
    if (slot < 8) return 7 + slot;
    if (slot == 8) return 23;
    if (slot == 9) return 24;
    if (slot == 10) return 27;
    if (slot == 11) return 131;
    if (slot < 15) return 324 + (slot - 12);
    if (slot < 18) return 519 + (slot - 15);
    if (slot < 21) return 714 + (slot - 18);
    if (slot < 27) return 717 + (slot - 21);
    return 835;
}

struct EncCtx {
    const TableState* S;
    Tables T;
    u64* bm;        // [OBS_BM_ROWS] column masks of the 1.0 cells, zero on entry
    float* sv;      // [OBS_N_SPECIAL][34] value rows, zero on entry
    int seat;       // perspective (absolute seat)
    bool kan_select;
    int lane;
    const u8* dora_factor;  // [34]
    unsigned parts;         // which feature groups to derive (bit k = part k of ENC_PART_*), ENC_ALL_PARTS = everything
};

// Feature groups = contiguous row ranges, so that different warps can derive them independently:
// part 0 rows [0,132) hand/scalars/dora/own pond, 1 rows [132,717) opponents' ponds, 2 rows [717,874) counters,
// pond overview, melds, status, 3 rows [874,889) the action block. {first bm row, first sv slot} per part (+ end).
constexpr int ENC_N_PARTS = 4;
constexpr unsigned ENC_ALL_PARTS = 15;
MJX_HD int enc_part_bm_begin(int part) { return part == 0 ? 0 : part == 1 ? 132 : part == 2 ? 717 : part == 3 ? 874 : OBS_BM_ROWS; }
MJX_HD int enc_part_sv_begin(int part) { return part == 0 ? 0 : part == 1 ? 12 : part == 2 ? 21 : OBS_N_SPECIAL; }

// Writers. ENC_OR / ENC_VAL / ENC_MAXV may be called by any lane (different lanes, different or equal cells);
// the *_L0 forms are for values every lane holds identically: lane 0 writes.
#ifdef MJX_HOST_EMUL
#define ENC_OR(e, row, mask) do { (e).bm[(row)] |= (u64)(mask); } while (0)
#define ENC_VAL(e, row, col, v) do { (e).sv[enc_special_slot(row) * 34 + (col)] = (v); } while (0)
// "latest writer wins" cells whose value grows with the writer's position: a max (the reference assigns in order)
#define ENC_MAXV(e, row, col, v) do { float& x_ = (e).sv[enc_special_slot(row) * 34 + (col)]; x_ = max(x_, (v)); } while (0)
#define ENC_VALS_L0(e, row, v) do { for (int c_ = 0; c_ < 34; c_++) ENC_VAL(e, row, c_, v); } while (0)
#define ENC_OR_L0(e, row, mask) ENC_OR(e, row, mask)
// positions 0..n-1 of a list, one per lane on the device
#define ENC_FOR_POS(e, i, n) for (int i = 0; i < (n); i++)
#else
#define ENC_OR(e, row, mask) atomicOr(reinterpret_cast<unsigned long long*>(&(e).bm[(row)]), (unsigned long long)(mask))
#define ENC_VAL(e, row, col, v) do { (e).sv[enc_special_slot(row) * 34 + (col)] = (v); } while (0)
#define ENC_MAXV(e, row, col, v) atomicMax(reinterpret_cast<int*>(&(e).sv[enc_special_slot(row) * 34 + (col)]), __float_as_int(v))
#define ENC_VALS_L0(e, row, v) do { ENC_VAL(e, row, (e).lane, v); if ((e).lane < 2) ENC_VAL(e, row, 32 + (e).lane, v); } while (0)
#define ENC_OR_L0(e, row, mask) do { if ((e).lane == 0) (e).bm[(row)] |= (u64)(mask); } while (0)
#define ENC_FOR_POS(e, i, n) for (int i0_ = 0, i = (e).lane; i0_ < (n); i0_ += 32, i += 32)
#endif
#define ENC_ONES_L0(e, row) ENC_OR_L0(e, row, OBS_FULL_ROW)
#define ENC_ONE_L0(e, row, col) ENC_OR_L0(e, row, 1ull << (col))

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

// obs_repr.rs:694-712 over an explicit tile list accessor (T_NONE entries are skipped): the k-th copy of a kind
// lights plane `row + k`; planes row+4.. mark the three red fives. Every lane runs the same scan (uniform values).
template <typename F>
MJX_D void enc_tile_set(EncCtx& e, int row, int n, F get) {
    u64 c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    u32 akam = 0;
    for (int i = 0; i < n; i++) {
        const int tile = get(i);
        if (tile == T_NONE) continue;
        if (is_aka(tile)) akam |= 1u << (tile - T_5MR);
        const u64 bit = 1ull << deaka(tile);
        c4 |= c3 & bit; c3 |= c2 & bit; c2 |= c1 & bit; c1 |= bit;
    }
    if (c1) ENC_OR_L0(e, row, c1);
    if (c2) ENC_OR_L0(e, row + 1, c2);
    if (c3) ENC_OR_L0(e, row + 2, c3);
    if (c4) ENC_OR_L0(e, row + 3, c4);
    for (int k = 0; k < 3; k++) if ((akam >> k) & 1) ENC_ONES_L0(e, row + 4 + k);
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

// Derives the compact form (bm, sv) of the version-4 rows 0..888 (889..1011, the single-player block, is
// k_sp_finalize's). Called by ONE warp; bm and sv must be zero on entry.
MJX_DN void encode_obs_v4(EncCtx& e, const Ctx& c, u64* mask_out) {
    const TableState* S = e.S;
    const int p = e.seat;
    const SeatPrivate& P = S->priv[p];
    const u16 cans = P.cans;
    const u8* df = e.dora_factor;

    // ---- hand (rows 0-6)
    if (e.parts & 1) {
        u64 m1, m2, m3, m4;
        tile_eval2(c, true, [&](int t) { const int n = P.tehai[t]; return (n > 0 ? 1 : 0) | (n > 1 ? 2 : 0); }, m1, m2);
        tile_eval2(c, true, [&](int t) { const int n = P.tehai[t]; return (n > 2 ? 1 : 0) | (n > 3 ? 2 : 0); }, m3, m4);
        ENC_OR_L0(e, 0, m1); ENC_OR_L0(e, 1, m2); ENC_OR_L0(e, 2, m3); ENC_OR_L0(e, 3, m4);
        for (int k = 0; k < 3; k++) if ((P.akas_in_hand >> k) & 1) ENC_ONES_L0(e, 4 + k);
    }
    // ---- scalars (rows 7-27)
    if (e.parts & 1) {
        int rank = 0;
        for (int i = 0; i < 4; i++) {
            i32 sc = S->scores[rel_to_abs(p, i)];
            ENC_VALS_L0(e, 7 + 2 * i, (float)min(max(sc, 0), 100000) / 100000.f);
            ENC_VALS_L0(e, 8 + 2 * i, (float)min(max(sc, 0), 30000) / 30000.f);
        }
        for (int s = 0; s < 4; s++)  // rankings.rs:8-22: stable by seat
            if (s != p && (S->scores[s] > S->scores[p] || (S->scores[s] == S->scores[p] && s < p))) rank++;
        ENC_ONES_L0(e, 15 + rank);
        const int kyoku_in_wind = S->kyoku & 3, bakaze = T_E + S->kyoku / 4;
        ENC_ONES_L0(e, 19 + kyoku_in_wind);
        ENC_VALS_L0(e, 23, (float)min((int)S->honba, 10) / 10.f);
        ENC_VALS_L0(e, 24, (float)min((int)S->kyotaku, 10) / 10.f);
        ENC_ONE_L0(e, 25, bakaze);
        ENC_ONE_L0(e, 26, T_E + ((p + 4 - S->oya) & 3));
        int gk = min(bakaze - T_E, 1) * 4 + kyoku_in_wind;
        ENC_VALS_L0(e, 27, (float)min(gk, 7) / 7.f);
    }
    // ---- dora indicators (rows 28-34)
    if (e.parts & 1) enc_tile_set(e, 28, S->n_dora, [&](int i) { return dora_indicator(S, i); });
    // ---- counters (rows 717-722)
    if (e.parts & 4) {
        ENC_VALS_L0(e, 717, (float)S->tiles_left / 69.f);
        // sum_t seen[t] * dora_factor[t] == sum over indicators of seen[indicated tile]
        int seen_doras = mjx_popc((u32)(S->akas_public | P.akas_in_hand));
        int own_doras = mjx_popc((u32)P.akas_in_hand);
        for (int k = 0; k < S->n_dora; k++) {
            const int d = tile_next(dora_indicator(S, k));
            seen_doras += S->public_seen[d] + P.tehai[d];
            own_doras += P.tehai[d];
        }
        for (int i = 0; i < 4; i++) {
            const int s = rel_to_abs(p, i);
            const SeatPublic& U = S->pub[s];
            int n = i == 0 ? own_doras : 0;
            for (int f = 0; f < U.n_fuuro; f++)
                for (int j = 0; j < 4; j++) {
                    int t = U.fuuro[f][j];
                    if (t != T_NONE) n += df[deaka(t)] + (is_aka(t) ? 1 : 0);
                }
            for (int j = 0; j < U.n_ankan; j++) {
                int t = U.ankan[j];
                n += 4 * df[t] + ((t == T_5M || t == T_5P || t == T_5S) ? 1 : 0);
            }
            ENC_VALS_L0(e, 718 + i, (float)min(n, 12) / 12.f);
        }
        int unseen = (S->n_dora * 4 + 3 - seen_doras) & 0xFF;
        ENC_VALS_L0(e, 722, (float)min(unseen, 23) / 23.f);
    }
    // max kawa length over the four ponds as this seat sees them (obs_repr.rs:221)
    int max_kawa_len = 0;
    for (int s = 0; s < 4; s++) max_kawa_len = max(max_kawa_len, kawa_view(S, p, s).len());

    // ---- own pond (rows 35-131): lanes are pond positions
    if (e.parts & 1) {
        const KawaView kv = kawa_view(S, p, p);
        const int len = kv.len();
        ENC_FOR_POS(e, i, len) {
            const KawaItem* k = i < len ? kv.at(i) : nullptr;
            if (!k) continue;
            const int tile = k->tile, fl = k->flags;
            u64 kans = 0;
            for (int q = 0; q < 4; q++) if (k->kan[q] != T_NONE) kans |= 1ull << k->kan[q];
            for (int pass = 0; pass < 2; pass++) {
                const int j = pass == 0 ? i : len - 1 - i;
                if (j >= (pass == 0 ? 6 : 18)) continue;
                const int row = (pass == 0 ? 35 : 59) + 4 * j;
                if (kans) ENC_OR(e, row, kans);
                ENC_OR(e, row + 1, 1ull << deaka(tile));
                if (is_aka(tile)) ENC_OR(e, row + 2, OBS_FULL_ROW);
                if (fl & SF_DORA) ENC_OR(e, row + 3, OBS_FULL_ROW);
            }
            ENC_MAXV(e, 131, deaka(tile), expf(-0.2f * (float)(max_kawa_len - 1 - i)));
        }
    }
    // ---- the three opponents' ponds (rows 132-716)
    if (e.parts & 2) for (int rel = 1; rel < 4; rel++) {
        const int sec = 132 + 195 * (rel - 1);
        const KawaView kv = kawa_view(S, p, rel_to_abs(p, rel));
        const int len = kv.len();
        ENC_FOR_POS(e, i, len) {
            const KawaItem* k = i < len ? kv.at(i) : nullptr;
            if (!k) continue;
            const int tile = k->tile, fl = k->flags;
            u64 kans = 0;
            for (int q = 0; q < 4; q++) if (k->kan[q] != T_NONE) kans |= 1ull << k->kan[q];
            for (int pass = 0; pass < 2; pass++) {
                const int j = pass == 0 ? i : len - 1 - i;
                if (j >= (pass == 0 ? 6 : 18)) continue;
                const int row = sec + (pass == 0 ? 0 : 48) + 8 * j;
                if (fl & SF_HAS_CHIPON) {
                    ENC_OR(e, row, 1ull << min(k->consumed[0], k->consumed[1]));
                    ENC_OR(e, row + 1, 1ull << max(k->consumed[0], k->consumed[1]));
                }
                if (kans) ENC_OR(e, row + 2, kans);
                ENC_OR(e, row + 3, 1ull << deaka(tile));
                if (is_aka(tile)) ENC_OR(e, row + 4, OBS_FULL_ROW);
                if (fl & SF_DORA) ENC_OR(e, row + 5, OBS_FULL_ROW);
                if (fl & SF_TEDASHI) ENC_OR(e, row + 6, OBS_FULL_ROW);
                if (fl & SF_RIICHI) ENC_OR(e, row + 7, OBS_FULL_ROW);
            }
            const float v = expf(-0.2f * (float)(max_kawa_len - 1 - i));
            const int tid = deaka(tile);
            ENC_MAXV(e, sec + 192, tid, v);
            if (fl & SF_TEDASHI) ENC_MAXV(e, sec + 193, tid, v);
            if (fl & SF_RIICHI) ENC_MAXV(e, sec + 194, tid, v);
        }
    }
    // ---- kawa overview (rows 723-750): real discards only (update.rs:336); empty slots hold T_NONE
    if (e.parts & 4) for (int i = 0; i < 4; i++) {
        const SeatPublic& U = S->pub[rel_to_abs(p, i)];
        enc_tile_set(e, 723 + 7 * i, U.kawa_len, [&](int j) { return (int)U.kawa[j].tile; });
    }
    // ---- melds (rows 751-834)
    if (e.parts & 4) for (int i = 0; i < 4; i++) {
        const SeatPublic& U = S->pub[rel_to_abs(p, i)];
        for (int f = 0; f < U.n_fuuro; f++) {
            const int row = 751 + 20 * i + 5 * f;
            u64 c1 = 0, c2 = 0, c3 = 0, c4 = 0;  // k-th copy of a kind goes to the first still-zero plane (obs_repr.rs:305-308)
            bool aka = false;
            for (int j = 0; j < 4; j++) {
                const int t = U.fuuro[f][j];
                if (t == T_NONE) continue;
                aka = aka || is_aka(t);
                const u64 bit = 1ull << deaka(t);
                c4 |= c3 & bit; c3 |= c2 & bit; c2 |= c1 & bit; c1 |= bit;
            }
            ENC_OR_L0(e, row, c1);
            if (c2) ENC_OR_L0(e, row + 1, c2);
            if (c3) ENC_OR_L0(e, row + 2, c3);
            if (c4) ENC_OR_L0(e, row + 3, c4);
            if (aka) ENC_ONES_L0(e, row + 4);
        }
        u64 ak = 0;
        for (int j = 0; j < U.n_ankan; j++) ak |= 1ull << U.ankan[j];
        if (ak) ENC_OR_L0(e, 831 + i, ak);
    }
    // ---- seen tiles, key discards, riichi / wait status (rows 835-873)
    if (e.parts & 4) {
        MJX_FOR_TILES(e, t) { ENC_VAL(e, 835, t, (float)(S->public_seen[t] + P.tehai[t]) / 4.f); }
        if (P.waits) ENC_OR_L0(e, 860, P.waits);
        for (int rel = 1; rel < 4; rel++) {
            const SeatPublic& U = S->pub[rel_to_abs(p, rel)];
            if (U.last_tedashi_flags & SF_VALID) {
                const int row = 836 + 3 * (rel - 1);
                ENC_ONE_L0(e, row, deaka(U.last_tedashi_tile));
                if (is_aka(U.last_tedashi_tile)) ENC_ONES_L0(e, row + 1);
                if (U.last_tedashi_flags & SF_DORA) ENC_ONES_L0(e, row + 2);
            }
            if (U.riichi_flags & SF_VALID) {
                const int row = 845 + 3 * (rel - 1);
                ENC_ONE_L0(e, row, deaka(U.riichi_tile));
                if (is_aka(U.riichi_tile)) ENC_ONES_L0(e, row + 1);
                if (U.riichi_flags & SF_DORA) ENC_ONES_L0(e, row + 2);
            }
            if ((S->riichi_declared >> rel_to_abs(p, rel)) & 1) ENC_ONES_L0(e, 854 + rel - 1);
            if ((S->riichi_accepted >> rel_to_abs(p, rel)) & 1) ENC_ONES_L0(e, 857 + rel - 1);
        }
        if (P.flags & PF_AT_FURITEN) ENC_ONES_L0(e, 861);
        ENC_ONES_L0(e, 862 + min(max((int)P.shanten, 0), 6));
        if ((S->riichi_accepted >> p) & 1) ENC_ONES_L0(e, 869);
        if (e.kan_select) ENC_ONES_L0(e, 870);
        if (cans & CAN_PASS) {
            const int tile = S->last_kawa_tile, tid = deaka(tile);
            ENC_ONE_L0(e, 871, tid);
            if (is_aka(tile)) ENC_ONES_L0(e, 872);
            if (df[tid] > 0) ENC_ONES_L0(e, 873);
        }
    }
    // ---- the action block (rows 874-888) + legal mask
    if (e.parts & 8) {
        const u64 discards = (cans & CAN_DISCARD) ? discard_candidates(c, p) : 0;  // warp collective
        if (mask_out) *mask_out = legal_mask(c, p, e.kan_select, discards);
        if (cans & CAN_DISCARD) {
            u64 d34 = (discards & ((1ull << 34) - 1)) | (((discards >> 34) & 1) << 4) | (((discards >> 35) & 1) << 13) |
                      (((discards >> 36) & 1) << 22);
            u64 ut = 0;
            if (P.shanten <= 1) ut = unconditional_tenpai_discards(e, c);
            ENC_OR_L0(e, 874, d34);
            ENC_OR_L0(e, 875, P.keep_shanten);
            ENC_OR_L0(e, 876, P.next_shanten);
            ENC_OR_L0(e, 877, ut);
            if ((S->riichi_declared >> p) & 1) ENC_ONES_L0(e, 878);
        }
        if (cans & CAN_RIICHI) ENC_ONES_L0(e, 879);
        if (cans & CAN_CHI_LOW) ENC_ONES_L0(e, 880);
        if (cans & CAN_CHI_MID) ENC_ONES_L0(e, 881);
        if (cans & CAN_CHI_HIGH) ENC_ONES_L0(e, 882);
        if (cans & CAN_PON) ENC_ONES_L0(e, 883);
        if (cans & CAN_DAIMINKAN) ENC_ONES_L0(e, 884);
        if (cans & CAN_ANKAN) ENC_OR_L0(e, 885, P.ankan_cand);
        if (cans & CAN_KAKAN) ENC_OR_L0(e, 886, P.kakan_cand);
        if (cans & CAN_AGARI) ENC_ONES_L0(e, 887);
        if (cans & CAN_RYUKYOKU) ENC_ONES_L0(e, 888);
    }
}

// Materialise obs rows [row_lo, row_hi) from the compact form into `tile` ((row_hi - row_lo) x 34 floats).
// Called by ONE warp after encode_obs_v4 (and a warp sync). On the device the caller passes the slice's row masks
// already in registers (lane l holds rows row_lo + l and row_lo + 32 + l), loaded ahead of time to hide their latency.
#ifdef MJX_HOST_EMUL
MJX_DN void enc_materialize(const EncCtx& e, float* tile, int row_lo, int row_hi) {
    for (int r = row_lo; r < row_hi; r++) {
        const u64 m = r < OBS_BM_ROWS ? e.bm[r] : 0;
        const int slot = enc_special_slot(r);
        for (int col = 0; col < 34; col++)
            tile[(r - row_lo) * 34 + col] = slot >= 0 ? e.sv[slot * 34 + col] : (((m >> col) & 1) ? 1.f : 0.f);
    }
}
#else
static_assert(OBS_SLICE_ROWS <= 64, "two mask words per lane cover a slice");
MJX_D void enc_load_masks(const u64* bm, int row_lo, int lane, u64& m0, u64& m1) {
    const int r0 = row_lo + lane, r1 = row_lo + 32 + lane;
    m0 = r0 < OBS_BM_ROWS ? __ldg(bm + r0) : 0;
    m1 = (32 + lane < OBS_SLICE_ROWS && r1 < OBS_BM_ROWS) ? __ldg(bm + r1) : 0;
}
MJX_DN void enc_materialize(const float* sv, int lane, float* tile, int row_lo, int row_hi, u64 m0, u64 m1) {
    {   // zero the slice: almost all of it stays zero
        uint4* t4 = reinterpret_cast<uint4*>(tile);
        const uint4 z = make_uint4(0, 0, 0, 0);
        const int n16 = (row_hi - row_lo) * 34 * 4 / 16;
        for (int i = lane; i < n16; i += 32) t4[i] = z;
    }
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const u64 m = half == 0 ? m0 : m1;
        unsigned nz = __ballot_sync(0xffffffffu, m != 0);
        while (nz) {  // one non-zero row at a time, lanes are columns
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)m, src);
            const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(m >> 32), src);
            float* dst = tile + (half * 32 + src) * 34;
            if ((lo >> lane) & 1) dst[lane] = 1.f;
            if (lane < 2 && ((hi >> lane) & 1)) dst[32 + lane] = 1.f;
        }
    }
    {   // value rows inside this slice: lane s looks at slot s, the hits are copied by the whole warp
        const int my_row = lane < OBS_N_SPECIAL ? enc_special_row(lane) : -1;
        unsigned hit = __ballot_sync(0xffffffffu, my_row >= row_lo && my_row < row_hi);
        while (hit) {
            const int slot = __ffs(hit) - 1;
            hit &= hit - 1;
            float* dst = tile + (__shfl_sync(0xffffffffu, my_row, slot) - row_lo) * 34;
            const float* src = sv + slot * 34;
            dst[lane] = __ldg(src + lane);
            if (lane < 2) dst[32 + lane] = __ldg(src + 32 + lane);
        }
    }
}
#endif

}  // namespace mjx
