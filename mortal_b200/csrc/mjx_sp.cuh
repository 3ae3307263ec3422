// mortal_b200 — single-player tables (obs v4 rows 889..1011) on device.
//
// Contract: libriichi state/agent_helper.rs:509-593 (single_player_tables) -> algo/sp/calc.rs
// (SPCalculator with calc_tegawari = calc_shanten_down = maximize_win_prob = false, which is how
// PlayerState calls it) -> state/obs_repr.rs:561-617, 632-692.
//
// The reference is a memoised recursion (AHashMap<State, Rc<Values>> per shanten level). Here the same
// quantities are computed as a level-synchronous dynamic programme by one CTA per observation:
//   1. expand: breadth-first over the DAG of reachable (hand, wall) states, alternating
//      W-states (3n+1, waiting for a useful draw) and D-states (3n+2, choosing a shanten-keeping discard);
//      one warp expands one state, lanes = the 34 tile ids (shanten of hand+-tile per lane), children are
//      de-duplicated through an open-addressing hash table in the CTA's global workspace;
//   2. evaluate: levels in reverse; one warp per state, lane i owns turn i and accumulates over the state's
//      edges in the reference's iteration order with explicitly rounded f32 ops, so every per-turn
//      tenpai / win / EV value follows the same sequence of roundings as the Rust code.
#pragma once
#include "mjx_obs.cuh"

namespace mjx {

constexpr int SP_T_MAX = 17;             // sp/mod.rs:42 MAX_TSUMOS_LEFT
constexpr int SP_SHANTEN_THRES = 3;      // calc.rs:13
constexpr int SP_MAX_TILES_LEFT = 34 * 4 - 1 - 13;  // calc.rs:14
constexpr int SP_EDGE_MAX = 40;          // <= 37 draw kinds / <= 14 discards
constexpr int SP_NODE_CAP = 32768;
constexpr int SP_HASH_CAP = 65536;
constexpr u32 SP_NO_CHILD = 0xFFFFFu;

#ifdef MJX_HOST_EMUL
#define SP_FMUL(a, b) ((a) * (b))
#define SP_FADD(a, b) ((a) + (b))
#define SP_FDIV(a, b) ((a) / (b))
#define SP_CTA_SYNC() ((void)0)
#else
// never contracted into FMA: the reference rounds after every multiply and add
#define SP_FMUL(a, b) __fmul_rn((a), (b))
#define SP_FADD(a, b) __fadd_rn((a), (b))
#define SP_FDIV(a, b) __fdiv_rn((a), (b))
#define SP_CTA_SYNC() __syncthreads()
#endif

// sp/state.rs:10-21 (n_extra_tsumo is always 0 without tegawari)
struct SpKey {
    u8 tehai[34];
    u8 wall[34];
    u8 akas;      // bits 0-2 akas_in_hand, bits 3-5 akas_in_wall
    u8 pad_[3];
};
static_assert(sizeof(SpKey) == 72, "SpKey layout");

struct SpWork {  // one per CTA, in global memory
    SpKey* keys;      // [SP_NODE_CAP]
    float* vals;      // [SP_NODE_CAP][3][SP_T_MAX]
    u32* edges;       // [SP_NODE_CAP][SP_EDGE_MAX]  child(20) | tile(6) << 20 | count(3) << 26
    u8* n_edges;      // [SP_NODE_CAP]
    u32* hash;        // [SP_HASH_CAP] node index + 1, 0 = empty
    i32* counters;    // [0] n_nodes, [1] overflow flag
};

struct SpParams {  // sp/calc.rs:36-62 + per-call arguments
    u8 tehai_len_div3;
    bool is_menzen, prefer_riichi, calc_double_riichi, calc_haitei;
    u8 bakaze, jikaze, num_doras_in_fuuro;
    u8 n_dora;
    u8 dora_ind[5];
    const u8 *chis, *pons, *minkans, *ankans;
    int n_chis, n_pons, n_minkans, n_ankans;
    int T;        // tsumos_left = MAX_TSUMO
    int n_left;   // tiles in the wall at the root
};

struct SpShared {  // per-CTA shared scratch
    float tsumo_prob[4][SP_T_MAX];                          // calc.rs:136-146
    float not_tsumo_prob[SP_MAX_TILES_LEFT + 1][SP_T_MAX];  // calc.rs:148-167
    float scores[8][SP_EDGE_MAX][4];                        // per warp: get_score of each edge of a W0 state
    u8 score_ok[8][SP_EDGE_MAX];
    i32 level_begin[10];
    i32 n_levels;
    SpParams P;
    u8 root_tehai[34], root_wall[34];
    u8 melds[16];
};

MJX_CONST float c_uradora_prob[5][13] = {  // algo/data/uradora_prob_table.txt
    {0.639485f, 0.327801f, 0.0327134f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.406736f, 0.42281f, 0.147966f, 0.021674f, 0.0008142f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.257516f, 0.406819f, 0.246851f, 0.0757724f, 0.0122266f, 0.0008004f, 1.43e-5f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.162199f, 0.346513f, 0.301539f, 0.142396f, 0.0401276f, 0.0066491f, 0.0005575f, 1.85e-5f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.101768f, 0.275319f, 0.313742f, 0.20189f, 0.081774f, 0.0215394f, 0.0035918f, 0.0003607f, 1.52e-5f, 3e-7f, 0.f, 0.f, 0.f},
};
MJX_CONST u8 c_discard_priority[38] = {6, 5, 4, 3, 2, 3, 4, 5, 6, 6, 5, 4, 3, 2, 3, 4, 5, 6, 6, 5, 4, 3, 2, 3, 4, 5, 6,
                                        7, 7, 7, 7, 7, 7, 7, 1, 1, 1, 0};  // tile.rs:20-27

// tile.rs:177-185
MJX_D int cmp_discard_priority(int l, int r) {
    int pl = c_discard_priority[l], pr = c_discard_priority[r];
    if (pl != pr) return pl < pr ? -1 : 1;
    if (l != r) return r < l ? -1 : 1;
    return 0;
}

struct SpCtx {
    SpWork W;
    SpShared* sh;
    Tables T;
    int lane, warp, nwarps;
};

MJX_D u32 sp_hash_key(const SpKey& k) {
    const u32* w = reinterpret_cast<const u32*>(&k);
    u32 h = 2166136261u;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SpKey) / 4); i++) { h ^= w[i]; h *= 16777619u; h ^= h >> 15; }
    return h;
}

MJX_D bool sp_key_eq(const SpKey& a, const SpKey& b) {
    const u32* x = reinterpret_cast<const u32*>(&a);
    const u32* y = reinterpret_cast<const u32*>(&b);
    bool eq = true;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SpKey) / 4); i++) eq &= x[i] == y[i];
    return eq;
}

// find-or-insert; executed by ONE lane. Returns node index, or -1 on overflow.
MJX_DN int sp_intern(SpCtx& s, const SpKey& key) {
    u32 slot = sp_hash_key(key) & (SP_HASH_CAP - 1);
    for (int probe = 0; probe < SP_HASH_CAP; probe++, slot = (slot + 1) & (SP_HASH_CAP - 1)) {
#ifdef MJX_HOST_EMUL
        u32 cur = s.W.hash[slot];
        if (cur == 0) {
            int idx = s.W.counters[0];
            if (idx >= SP_NODE_CAP) { s.W.counters[1] = 1; return -1; }
            s.W.counters[0] = idx + 1;
            s.W.keys[idx] = key;
            s.W.n_edges[idx] = 0;
            s.W.hash[slot] = (u32)idx + 1;
            return idx;
        }
        if (sp_key_eq(s.W.keys[cur - 1], key)) return (int)cur - 1;
#else
        u32 cur = atomicAdd(&s.W.hash[slot], 0u);
        if (cur == 0) {
            // claim the slot with a sentinel, publish the key, then the index
            u32 prev = atomicCAS(&s.W.hash[slot], 0u, 0xFFFFFFFFu);
            if (prev == 0) {
                int idx = atomicAdd(&s.W.counters[0], 1);
                if (idx >= SP_NODE_CAP) { s.W.counters[1] = 1; atomicExch(&s.W.hash[slot], 0u); return -1; }
                s.W.keys[idx] = key;
                s.W.n_edges[idx] = 0;
                __threadfence_block();
                atomicExch(&s.W.hash[slot], (u32)idx + 1);
                return idx;
            }
            cur = prev;
        }
        while (cur == 0xFFFFFFFFu) cur = atomicAdd(&s.W.hash[slot], 0u);  // another warp is publishing
        __threadfence_block();
        if (sp_key_eq(s.W.keys[cur - 1], key)) return (int)cur - 1;
#endif
    }
    s.W.counters[1] = 1;
    return -1;
}

// Expand one node (one warp). is_w: W-state at shanten k (edges = useful draws) else D-state at shanten k
// (edges = shanten-keeping discards). Children are interned unless `leaf`.
MJX_DN void sp_expand(SpCtx& s, const Ctx& c, int node, bool is_w, int k, bool leaf) {
    const SpKey key = s.W.keys[node];
    const int len = s.sh->P.tehai_len_div3;
    const HandSig base = hand_sig(key.tehai);
    u64 eff, unused;
    if (is_w) {
        tile_eval2(c, true, [&](int t) {
            if (key.wall[t] == 0) return 0;
            return shanten_all_sig(s.T, sig_variant(base, t, +1, key.tehai[t]), len) - k == -1 ? 1 : 0;
        }, eff, unused);
    } else {
        tile_eval2(c, true, [&](int t) {
            if (key.tehai[t] == 0) return 0;
            return shanten_all_sig(s.T, sig_variant(base, t, -1, key.tehai[t]), len) == k ? 1 : 0;
        }, eff, unused);
    }
    if (MJX_IS_L0(c)) {
        int ne = 0;
        for (u64 rest = eff; rest; rest &= rest - 1) {
            const int t = mjx_ffsll(rest) - 1;
            const int suit5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
            if (is_w) {
                const int count = key.wall[t];
                const bool aka_in_wall = suit5 >= 0 && ((key.akas >> (3 + suit5)) & 1);
                for (int variant = 0; variant < 2; variant++) {
                    int tile, cnt;
                    if (aka_in_wall) {
                        if (variant == 0) { if (count < 2) continue; tile = t; cnt = count - 1; }
                        else { tile = T_5MR + suit5; cnt = 1; }
                    } else {
                        if (variant == 1) break;
                        tile = t; cnt = count;
                    }
                    u32 child = SP_NO_CHILD;
                    if (!leaf) {
                        SpKey ck = key;
                        ck.tehai[t] += 1;
                        ck.wall[t] -= 1;
                        if (is_aka(tile)) ck.akas = (u8)((ck.akas | (1 << suit5)) & ~(1 << (3 + suit5)));
                        int ci = sp_intern(s, ck);
                        if (ci < 0) break;
                        child = (u32)ci;
                    }
                    if (ne < SP_EDGE_MAX) s.W.edges[(size_t)node * SP_EDGE_MAX + ne++] = child | ((u32)tile << 20) | ((u32)cnt << 26);
                }
            } else {
                // sp/state.rs:127-132: the aka is discarded only when it is the last 5 of its suit in hand
                int tile = t;
                if (suit5 >= 0 && ((key.akas >> suit5) & 1) && key.tehai[t] == 1) tile = T_5MR + suit5;
                SpKey ck = key;
                ck.tehai[t] -= 1;
                if (is_aka(tile)) ck.akas = (u8)(ck.akas & ~(1 << suit5));
                int ci = sp_intern(s, ck);
                if (ci < 0) break;
                if (ne < SP_EDGE_MAX) s.W.edges[(size_t)node * SP_EDGE_MAX + ne++] = (u32)ci | ((u32)tile << 20);
            }
        }
        s.W.n_edges[node] = (u8)ne;
    }
    MJX_SYNCWARP();
}

// calc.rs:640-758 for one winning draw; executed by one lane. Returns false when there is no yaku.
MJX_DN bool sp_get_score(const SpCtx& s, const SpKey& key, int win_tile, float* scores) {
    const SpParams& P = s.sh->P;
    u8 th[34];
    for (int i = 0; i < 34; i++) th[i] = key.tehai[i];
    const int wid = deaka(win_tile);
    th[wid] += 1;
    const int akas_in_hand = (key.akas & 7) | (is_aka(win_tile) ? (1 << (win_tile - T_5MR)) : 0);
    u8 wall[34];
    for (int i = 0; i < 34; i++) wall[i] = key.wall[i];
    wall[wid] -= 1;
    AgariQuery q;
    q.tehai = th;
    q.chis = P.chis; q.pons = P.pons; q.minkans = P.minkans; q.ankans = P.ankans;
    q.n_chis = P.n_chis; q.n_pons = P.n_pons; q.n_minkans = P.n_minkans; q.n_ankans = P.n_ankans;
    q.bakaze = P.bakaze; q.jikaze = P.jikaze; q.winning_tile = wid; q.is_ron = false; q.is_menzen = P.is_menzen;
    const bool is_oya = P.jikaze == T_E;
    const int additional = P.is_menzen ? (P.prefer_riichi ? 2 : 1) : 0;
    int doras = mjx_popc((u32)akas_in_hand) + P.num_doras_in_fuuro;
    for (int i = 0; i < P.n_dora; i++) doras += th[tile_next(P.dora_ind[i])];
    Agari a = agari_with(s.T, q, additional, doras & 0xFF);
    if (a.kind == 0) return false;
    if (a.kind == 2) {
        float v = (float)tsumo_total(point_yakuman(is_oya, a.yakuman), is_oya);
        for (int i = 0; i < 4; i++) scores[i] = v;
        return true;
    }
    const int fu = a.fu, han = a.han;
    auto pts = [&](int h) { bool ok; return (float)tsumo_total(point_calc(is_oya, fu, h, &ok), is_oya); };
    const bool assume_riichi = P.is_menzen && P.prefer_riichi;
    for (int i = 0; i < 4; i++) scores[i] = 0.f;
    if (assume_riichi && P.n_dora == 1) {
        int n_ind[5] = {0, 0, 0, 0, 0};
        int sum_ind = 0, n_left = 0;
        for (int t = 0; t < 34; t++) n_left += wall[t];
        for (int t = 0; t < 34; t++) {
            int cnt = th[t];
            if (cnt == 0) continue;
            int ic = wall[tile_prev(t)];
            n_ind[cnt] = (n_ind[cnt] + ic) & 0xFF;
            sum_ind = (sum_ind + ic) & 0xFF;
        }
        float up[5];
        up[0] = SP_FDIV((float)((n_left - sum_ind) & 0xFF), (float)n_left);
        for (int i = 1; i < 5; i++) up[i] = SP_FDIV((float)n_ind[i], (float)n_left);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 5; j++) {
                if (up[j] == 0.f) continue;
                scores[i] = SP_FADD(scores[i], SP_FMUL(pts(han + i + j), up[j]));
            }
    } else if (assume_riichi && P.n_dora > 1) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 13; j++) {
                float p = c_uradora_prob[P.n_dora - 1][j];
                if (p == 0.f) continue;
                scores[i] = SP_FADD(scores[i], SP_FMUL(pts(han + i + j), p));
            }
    } else {
        for (int i = 0; i < 4; i++) scores[i] = pts(han + i);
    }
    return true;
}

MJX_D float* sp_vals(const SpCtx& s, int node, int which) { return s.W.vals + ((size_t)node * 3 + which) * SP_T_MAX; }

// calc.rs:447-561 draw_without_tegawari_slow for one W-state at shanten k (one warp, lane i = turn i)
MJX_DN void sp_eval_w(SpCtx& s, const Ctx& c, int node, int k) {
    const SpParams& P = s.sh->P;
    const int T = P.T;
    const int ne = s.W.n_edges[node];
    const u32* edges = s.W.edges + (size_t)node * SP_EDGE_MAX;
    int sum_required = 0;
    for (int e = 0; e < ne; e++) sum_required += (edges[e] >> 26) & 7;
    sum_required &= 0xFF;
    const float* not_tsumo = s.sh->not_tsumo_prob[sum_required <= SP_MAX_TILES_LEFT ? sum_required : SP_MAX_TILES_LEFT];
    float (*sc)[4] = s.sh->scores[s.warp];
    u8* sc_ok = s.sh->score_ok[s.warp];
    if (k == 0) {
        const SpKey key = s.W.keys[node];
#ifdef MJX_HOST_EMUL
        for (int e = 0; e < ne; e++) sc_ok[e] = sp_get_score(s, key, (edges[e] >> 20) & 63, sc[e]) ? 1 : 0;
#else
        for (int e = c.lane; e < ne; e += 32) sc_ok[e] = sp_get_score(s, key, (edges[e] >> 20) & 63, sc[e]) ? 1 : 0;
        __syncwarp();
#endif
    }
#ifdef MJX_HOST_EMUL
    for (int i = 0; i < T; i++) {
#else
    { const int i = c.lane; if (i < T) {
#endif
        float tenpai = 0.f, win = 0.f, ev = 0.f;
        const float m = not_tsumo[i];
        if (m != 0.f) {
            for (int e = 0; e < ne; e++) {
                const u32 ed = edges[e];
                const int cnt = (ed >> 26) & 7;
                if (k == 0 && !sc_ok[e]) continue;
                const float* tsumo_probs = s.sh->tsumo_prob[cnt - 1];
                const float *nt = nullptr, *nw = nullptr, *nv = nullptr;
                if (k > 0) {
                    const int child = (int)(ed & 0xFFFFF);
                    nt = sp_vals(s, child, 0); nw = sp_vals(s, child, 1); nv = sp_vals(s, child, 2);
                }
                for (int j = i; j < T; j++) {
                    const float n = not_tsumo[j];
                    if (n == 0.f) break;
                    const float prob = SP_FDIV(SP_FMUL(tsumo_probs[j], n), m);
                    if (k == 0) {
                        const bool assume_riichi = P.is_menzen && P.prefer_riichi;
                        const int han_plus = (assume_riichi && P.calc_double_riichi && i == 0) + (assume_riichi && j == i) +
                                             (P.calc_haitei && j == T - 1);
                        win = SP_FADD(win, prob);
                        ev = SP_FADD(ev, SP_FMUL(prob, sc[e][han_plus]));
                    } else {
                        if (k == 1) tenpai = SP_FADD(tenpai, prob);
                        if (j < T - 1) {
                            if (k > 1) tenpai = SP_FADD(tenpai, SP_FMUL(prob, nt[j + 1]));
                            win = SP_FADD(win, SP_FMUL(prob, nw[j + 1]));
                            ev = SP_FADD(ev, SP_FMUL(prob, nv[j + 1]));
                        }
                    }
                }
            }
        }
        sp_vals(s, node, 0)[i] = tenpai;
        sp_vals(s, node, 1)[i] = win;
        sp_vals(s, node, 2)[i] = ev;
#ifdef MJX_HOST_EMUL
    }
#else
    } }
    __syncwarp();
#endif
}

// calc.rs:563-637 discard_slow for one D-state (one warp, lane i = turn i)
MJX_DN void sp_eval_d(SpCtx& s, const Ctx& c, int node) {
    const int T = s.sh->P.T;
    const int ne = s.W.n_edges[node];
    const u32* edges = s.W.edges + (size_t)node * SP_EDGE_MAX;
#ifdef MJX_HOST_EMUL
    for (int i = 0; i < T; i++) {
#else
    { const int i = c.lane; if (i < T) {
#endif
        const float FMIN = -3.40282347e+38f;
        float bt = FMIN, bw = FMIN, bv = FMIN;
        int best_tile = T_UNK;
        i32 best_value = (i32)0x80000000;
        for (int e = 0; e < ne; e++) {
            const int child = (int)(edges[e] & 0xFFFFF), tile = (edges[e] >> 20) & 63;
            const float v = sp_vals(s, child, 2)[i];
            const i32 value = (i32)v;  // exp_values are finite and < 2^31 here; Rust `as i32` truncates the same way
            if (value > best_value || (value == best_value && cmp_discard_priority(tile, best_tile) > 0)) {
                bt = sp_vals(s, child, 0)[i]; bw = sp_vals(s, child, 1)[i]; bv = v;
                best_value = value; best_tile = tile;
            }
        }
        sp_vals(s, node, 0)[i] = bt;
        sp_vals(s, node, 1)[i] = bw;
        sp_vals(s, node, 2)[i] = bv;
#ifdef MJX_HOST_EMUL
    }
#else
    } }
    __syncwarp();
#endif
}

// per-candidate summary used by the obs rows
struct SpCand {
    int tile;           // as sp/candidate.rs (may be an aka id)
    int node;           // W-state whose values are the candidate's, or -1 (simple mode)
    u64 required;       // 34-bit set of required tile ids
    int num_required;   // sum of counts (u8 arithmetic in the reference)
    bool shanten_down;
    float t0, w0, e0;   // clamped first-turn values (comparators)
};

MJX_D float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
MJX_D int cmp_f32(float a, float b) { return a < b ? -1 : (a > b ? 1 : 0); }

// sp/candidate.rs:73-107
MJX_D int sp_cand_cmp(const SpCand& l, const SpCand& r, int by /*0 EV, 3 NotShantenDown*/, bool has_values) {
    if (l.tile == r.tile) return 0;
    if (by == 0 && has_values) {
        int o = cmp_f32(l.e0, r.e0); if (o) return o;
        o = cmp_f32(l.w0, r.w0); if (o) return o;
        o = cmp_f32(l.t0, r.t0); if (o) return o;
    }
    if (!l.shanten_down && r.shanten_down) return 1;
    if (l.shanten_down && !r.shanten_down) return -1;
    if (l.num_required != r.num_required) return l.num_required < r.num_required ? -1 : 1;
    return cmp_discard_priority(l.tile, r.tile);
}

// required tiles of a W-state hand (sp/state.rs:181-201): tiles in the wall that lower the shanten
MJX_DN void sp_required(const SpCtx& s, const Ctx& c, const u8* tehai, const u8* wall, u64* set, int* num) {
    const int len = s.sh->P.tehai_len_div3;
    const HandSig base = hand_sig(tehai);
    const int cur = shanten_all_sig(s.T, base, len);
    u64 req, unused;
    tile_eval2(c, true, [&](int t) {
        if (wall[t] == 0) return 0;
        return shanten_all_sig(s.T, sig_variant(base, t, +1, tehai[t]), len) < cur ? 1 : 0;
    }, req, unused);
    int n = 0;
    for (u64 rest = req; rest; rest &= rest - 1) n += wall[mjx_ffsll(rest) - 1];
    *set = req;
    *num = n & 0xFF;
}

// agent_helper.rs:467-503 from the table record
MJX_D int real_time_shanten(const Ctx& c, const TableState* S, int p) {
    const SeatPrivate& P = S->priv[p];
    if (!(P.cans & CAN_DISCARD)) return P.shanten;
    if (P.shanten > 0) return (P.flags & PF_HAS_NEXT_SHANTEN_DISCARD) ? P.shanten - 1 : P.shanten;
    if (P.last_self_tsumo != T_NONE) return ((P.waits >> deaka(P.last_self_tsumo)) & 1) ? -1 : 0;
    return shanten_all(c.T, P.tehai, P.tehai_len_div3);
}

// Encodes obs v4 rows 889..1011 into the (second-half) tile. Called by every thread of the CTA.
MJX_DN void encode_sp_block(EncCtx& e, const Ctx& c, SpCtx& s) {
    const TableState* S = e.S;
    const int p = e.seat;
    const SeatPrivate& PV = S->priv[p];
    const SeatPublic& PU = S->pub[p];
    const u16 cans = PV.cans;
    const u8* df = e.dora_factor;
    const bool tid0 = s.warp == 0 && s.lane == 0;

    // ---- agent_helper.rs:509-531: availability
    bool can_discard = (cans & CAN_DISCARD) != 0;
    const int cur_shanten = real_time_shanten(c, S, p);
    int tsumos_left = 0;
    bool calc_haitei = false;
    bool avail = S->tiles_left >= 4 && cur_shanten >= 0;
    if (avail) {
        if (can_discard) { tsumos_left = S->tiles_left / 4; calc_haitei = (S->tiles_left & 3) == 0; }
        else {
            int target = (PV.target_actor - p) & 3;
            int at_next = max((int)S->tiles_left - (4 - target), 0);
            tsumos_left = at_next / 4; calc_haitei = (at_next & 3) == 0;
        }
        avail = tsumos_left >= 1;
    }
    if (!avail) {
        // obs_repr.rs:604-616: min tsumo-agari points as the max EV, everything else skipped
        float v = 0.f;
        if (cans & CAN_AGARI) {
            bool ok;
            const bool is_ron = (cans & CAN_RON_AGARI) != 0;
            Point pt = agari_points(c, p, is_ron, 0, &ok);
            if (ok) v = (float)tsumo_total(pt, p == S->oya);
        }
        if (ENC_SECTION(e, 6, false)) {
            ENC_FILL(e, 889, fminf(fmaxf(v, 0.f), 100000.f) / 100000.f);
            ENC_FILL(e, 890, fminf(fmaxf(v, 0.f), 30000.f) / 30000.f);
        }
        return;
    }

    // ---- parameters (agent_helper.rs:533-585)
    SpShared* sh = s.sh;
    SP_CTA_SYNC();
    if (tid0) {
        SpParams& P = sh->P;
        P.tehai_len_div3 = PV.tehai_len_div3;
        P.is_menzen = (PV.flags & PF_IS_MENZEN) != 0;
        P.prefer_riichi = S->scores[p] >= 1000;
        P.calc_double_riichi = can_discard && (PV.flags & PF_CAN_W_RIICHI);
        P.calc_haitei = calc_haitei;
        P.bakaze = T_E + S->kyoku / 4;
        P.jikaze = T_E + ((p + 4 - S->oya) & 3);
        P.n_dora = S->n_dora;
        for (int i = 0; i < 5; i++) P.dora_ind[i] = i < S->n_dora ? (u8)dora_indicator(S, i) : 0;
        for (int i = 0; i < 4; i++) {
            sh->melds[i] = PV.chis[i]; sh->melds[4 + i] = PV.pons[i]; sh->melds[8 + i] = PV.minkans[i]; sh->melds[12 + i] = PV.ankans[i];
        }
        P.chis = sh->melds; P.pons = sh->melds + 4; P.minkans = sh->melds + 8; P.ankans = sh->melds + 12;
        P.n_chis = PV.n_chis; P.n_pons = PV.n_pons; P.n_minkans = PV.n_minkans; P.n_ankans = PV.n_ankans;
        // num_doras_in_fuuro = doras_owned[0] - doras in tehai - akas in hand (agent_helper.rs:533-545)
        int nf = 0;
        if (!(P.is_menzen && PU.n_ankan == 0)) {
            for (int f = 0; f < PU.n_fuuro; f++)
                for (int j = 0; j < 4; j++) { int t = PU.fuuro[f][j]; if (t != T_NONE) nf += df[deaka(t)] + (is_aka(t) ? 1 : 0); }
            for (int j = 0; j < PU.n_ankan; j++) { int t = PU.ankan[j]; nf += 4 * df[t] + ((t == T_5M || t == T_5P || t == T_5S) ? 1 : 0); }
        }
        P.num_doras_in_fuuro = (u8)nf;
        P.T = tsumos_left;
        // root hand / wall (InitState -> State, sp/state.rs:35-54)
        int akas_hand = PV.akas_in_hand;
        for (int t = 0; t < 34; t++) sh->root_tehai[t] = PV.tehai[t];
        const bool after_riichi = can_discard && ((S->riichi_accepted >> p) & 1);
        if (after_riichi) {
            int lt = PV.last_self_tsumo;
            sh->root_tehai[deaka(lt)] -= 1;
            if (is_aka(lt)) akas_hand &= ~(1 << (lt - T_5MR));
        }
        int n_left = 0;
        for (int t = 0; t < 34; t++) {
            int seen = S->public_seen[t] + PV.tehai[t];  // tiles_seen is NOT adjusted for the riichi discard
            sh->root_wall[t] = (u8)(4 - seen);
            n_left += 4 - seen;
        }
        P.n_left = n_left;
        SpKey root;
        for (int t = 0; t < 34; t++) { root.tehai[t] = sh->root_tehai[t]; root.wall[t] = sh->root_wall[t]; }
        const int akas_seen = S->akas_public | PV.akas_in_hand;
        root.akas = (u8)((akas_hand & 7) | (((~akas_seen) & 7) << 3));
        root.pad_[0] = root.pad_[1] = root.pad_[2] = 0;
        s.W.counters[0] = 1;
        s.W.counters[1] = 0;
        s.W.keys[0] = root;
        s.W.n_edges[0] = 0;
    }
    SP_CTA_SYNC();
    const bool after_riichi = can_discard && ((S->riichi_accepted >> p) & 1);
    if (after_riichi) can_discard = false;
    const SpParams& P = sh->P;
    const int T = P.T;

    // ---- candidate list
    SpCand cands[14];
    int n_cands = 0;
    const bool has_values = cur_shanten <= SP_SHANTEN_THRES;

    if (!has_values) {
        // calc.rs:281-314 analyze_*_simple: required tiles only; done redundantly by every warp (uniform)
        if (can_discard) {
            const HandSig base = hand_sig(sh->root_tehai);
            for (int t = 0; t < 34; t++) {
                if (sh->root_tehai[t] == 0) continue;
                u8 th[34];
                for (int i = 0; i < 34; i++) th[i] = sh->root_tehai[i];
                th[t] -= 1;
                int after = shanten_all_sig(s.T, sig_variant(base, t, -1, sh->root_tehai[t]), P.tehai_len_div3);
                SpCand& cd = cands[n_cands++];
                const int k5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
                cd.tile = (k5 >= 0 && ((s.W.keys[0].akas >> k5) & 1) && sh->root_tehai[t] == 1) ? T_5MR + k5 : t;
                cd.node = -1;
                cd.shanten_down = after - cur_shanten == 1;
                cd.t0 = cd.w0 = cd.e0 = 0.f;
                sp_required(s, c, th, sh->root_wall, &cd.required, &cd.num_required);
            }
        } else {
            SpCand& cd = cands[n_cands++];
            cd.tile = T_UNK; cd.node = -1; cd.shanten_down = false; cd.t0 = cd.w0 = cd.e0 = 0.f;
            sp_required(s, c, sh->root_tehai, sh->root_wall, &cd.required, &cd.num_required);
        }
    } else {
        // ---- probability tables (calc.rs:136-167), one row per thread
        {
#ifdef MJX_HOST_EMUL
            const int tid = 0, nthreads = 1;
#else
            const int tid = s.warp * 32 + s.lane, nthreads = s.nwarps * 32;
#endif
            const int n_left = P.n_left;
            for (int r = tid; r < 4 + SP_MAX_TILES_LEFT + 1; r += nthreads) {
                if (r < 4) {
                    for (int j = 0; j < T; j++) sh->tsumo_prob[r][j] = SP_FDIV((float)(r + 1), (float)(n_left - j));
                } else {
                    const int i = r - 4;
                    float* row = sh->not_tsumo_prob[i];
                    for (int j = 0; j < T; j++) row[j] = 0.f;
                    if (i <= n_left) {
                        row[0] = 1.f;
                        const int lim = min(T - 1, n_left - i);
                        for (int j = 0; j < lim; j++)
                            row[j + 1] = SP_FDIV(SP_FMUL(row[j], (float)(n_left - i - j)), (float)(n_left - j));
                    }
                }
            }
            // clear the hash table
            for (int i = tid; i < SP_HASH_CAP; i += nthreads) s.W.hash[i] = 0;
        }
        SP_CTA_SYNC();

        // ---- expand, level by level. Level sequence: [D_s root]? W_s D_{s-1} W_{s-1} ... D_0 W_0
        // level L kinds/shanten are derived from the root kind.
        const bool root_is_d = can_discard;
        if (tid0) { sh->level_begin[0] = 0; sh->level_begin[1] = 1; sh->n_levels = 1; }
        SP_CTA_SYNC();
        {
            int lvl = 0;
            bool is_w = !root_is_d;
            int k = cur_shanten;
            for (;;) {
                const int b = sh->level_begin[lvl], en = sh->level_begin[lvl + 1];
                const bool leaf = is_w && k == 0;
                for (int node = b + s.warp; node < en; node += s.nwarps) sp_expand(s, c, node, is_w, k, leaf);
                SP_CTA_SYNC();
                if (tid0) { sh->level_begin[lvl + 2] = min(s.W.counters[0], SP_NODE_CAP); sh->n_levels = lvl + 1; }
                SP_CTA_SYNC();
                if (leaf || s.W.counters[1]) break;
                // next level
                if (is_w) { is_w = false; k -= 1; } else { is_w = true; }
                lvl += 1;
            }
        }
        const bool overflow = s.W.counters[1] != 0;
        if (!overflow) {
            // ---- evaluate bottom-up
            const int n_levels = sh->n_levels;
            for (int lvl = n_levels - 1; lvl >= (root_is_d ? 1 : 0); lvl--) {
                // level kind: going down from the root the kinds alternate starting with root kind
                const bool is_w = root_is_d ? (lvl & 1) == 1 : (lvl & 1) == 0;
                const int k = root_is_d ? cur_shanten - lvl / 2 : cur_shanten - (lvl + 1) / 2;
                const int b = sh->level_begin[lvl], en = sh->level_begin[lvl + 1];
                for (int node = b + s.warp; node < en; node += s.nwarps) {
                    if (is_w) sp_eval_w(s, c, node, k); else sp_eval_d(s, c, node);
                }
                SP_CTA_SYNC();
            }
            // ---- candidates (calc.rs:203-279)
            if (root_is_d) {
                const int ne = s.W.n_edges[0];
                for (int i = 0; i < ne && n_cands < 14; i++) {
                    const u32 ed = s.W.edges[i];
                    SpCand& cd = cands[n_cands++];
                    cd.tile = (ed >> 20) & 63;
                    cd.node = (int)(ed & 0xFFFFF);
                    cd.shanten_down = false;
                }
            } else {
                SpCand& cd = cands[n_cands++];
                cd.tile = T_UNK; cd.node = 0; cd.shanten_down = false;
            }
            for (int i = 0; i < n_cands; i++) {
                SpCand& cd = cands[i];
                const int node = cd.node;
                const int ne = s.W.n_edges[node];
                u64 req = 0; int num = 0;
                for (int q = 0; q < ne; q++) {
                    const u32 ed = s.W.edges[(size_t)node * SP_EDGE_MAX + q];
                    req |= 1ull << deaka((ed >> 20) & 63);
                    num += (ed >> 26) & 7;
                }
                cd.required = req;
                cd.num_required = num & 0xFF;
                cd.t0 = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s, node, 0)[0]);
                cd.w0 = clamp01(sp_vals(s, node, 1)[0]);
                cd.e0 = fmaxf(sp_vals(s, node, 2)[0], 0.f);
            }
        }
        if (overflow) {
            if (tid0) { /* leave the block zero; the host reads the flag through mjx_env_sp_overflows */ }
            return;
        }
    }
    if (n_cands == 0) {
        // analyze_discard can return no candidate only if no discard keeps the shanten — impossible by construction
        return;
    }
    if (after_riichi) cands[0].tile = PV.last_self_tsumo;  // agent_helper.rs:588-590 (after sorting: see below)

    // ---- obs rows (obs_repr.rs:561-603, 644-692); written by one warp of the second half
    if (!ENC_SECTION(e, 6, false)) return;
    // `max_ev_table` is sorted descending by EV (stable); index 0 = maximum under sp_cand_cmp(…, EV)
    int first = 0;
    for (int i = 1; i < n_cands; i++) if (sp_cand_cmp(cands[i], cands[first], has_values ? 0 : 3, has_values) > 0) first = i;
    const float max_ev = has_values ? cands[first].e0 : 0.f;
    ENC_FILL(e, 889, fminf(fmaxf(max_ev, 0.f), 100000.f) / 100000.f);
    ENC_FILL(e, 890, fminf(fmaxf(max_ev, 0.f), 30000.f) / 30000.f);
    const bool cd_flag = (cans & CAN_DISCARD) != 0;  // obs_repr.rs uses cans.can_discard, not the riichi-adjusted flag
    if (cd_flag) {
        for (int i = 0; i < n_cands; i++) {
            const int dt = deaka(cands[i].tile);
            const int row = 891 + (cands[i].shanten_down ? 34 : 0) + dt;
            MJX_FOR_TILES(e, t) { if ((cands[i].required >> t) & 1) ENC_AT(e, row, t) = 1.f; }
        }
        // max_by(NotShantenDown) returns the LAST maximum
        int best = 0;
        for (int i = 1; i < n_cands; i++) if (sp_cand_cmp(cands[i], cands[best], 3, false) >= 0) best = i;
        ENC_ASSIGN(e, 959, deaka(cands[best].tile), 1.f);
    } else {
        MJX_FOR_TILES(e, t) { if ((cands[first].required >> t) & 1) ENC_AT(e, 960, t) = 1.f; }
    }
    if (!has_values) return;
    if (!(cands[first].t0 > 0.f)) return;  // obs_repr.rs:645-653
    const float ev_scale = max_ev < 1.f ? 0.f : SP_FDIV(1.f, max_ev);
    const int n_emit = cd_flag ? n_cands : 1;
    for (int q = 0; q < n_emit; q++) {
        const SpCand& cd = cd_flag ? cands[q] : cands[first];
        const int node = cd.node;
        for (int turn = 0; turn < T; turn++) {
            const float tp = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s, node, 0)[turn]);
            if (!(tp > 0.f)) break;
            const float wp = clamp01(sp_vals(s, node, 1)[turn]);
            const float ev = fminf(SP_FMUL(fmaxf(sp_vals(s, node, 2)[turn], 0.f), ev_scale), 1.f);
            if (cd_flag) {
                const int tid = deaka(cd.tile);
                ENC_ASSIGN(e, 961 + turn, tid, tp);
                ENC_ASSIGN(e, 961 + SP_T_MAX + turn, tid, wp);
                ENC_ASSIGN(e, 961 + 2 * SP_T_MAX + turn, tid, ev);
            } else {
                ENC_FILL(e, 961 + turn, tp);
                ENC_FILL(e, 961 + SP_T_MAX + turn, wp);
                ENC_FILL(e, 961 + 2 * SP_T_MAX + turn, ev);
            }
        }
    }
}

}  // namespace mjx
