// mortal_b200 — single-player tables (obs v4 rows 889..1011) on device.
//
// Contract: libriichi state/agent_helper.rs:509-593 (single_player_tables) -> algo/sp/calc.rs
// (SPCalculator with calc_tegawari = calc_shanten_down = maximize_win_prob = false, which is how
// PlayerState calls it) -> state/obs_repr.rs:561-617, 632-692.
//
// The reference is a memoised recursion per observation (AHashMap<State, Rc<Values>> per shanten level).
// Here ALL observations of a step are solved together as one level-synchronous dynamic programme over the
// union of their (hand, wall) state DAGs:
//   levels     D3 W3 D2 W2 D1 W1 D0 W0   (D = 3n+2 hand choosing a shanten-keeping discard,
//                                          W = 3n+1 hand waiting for a shanten-lowering draw)
//   state      a state of row r is the root hand plus the multiset of tiles drawn and the multiset discarded since
//              (sp/state.rs:10-21: tehai / wall / akas follow from those) — at most 3 draws and 4 discards, so the
//              memo key is ONE 64-bit word: row << 42 | discards (4 x 6 bit, sorted) << 18 | draws (3 x 6 bit, sorted).
//              States are interned by a single 64-bit CAS into an open-addressing table; the table slot is the state's
//              id and indexes its signature, edge range and value vectors — no key arrays, no second phase, no fences.
//   expand     one launch per level; a CTA takes a batch of 32 states: (state, tile) candidate pairs are compacted in shared
//              memory and evaluated one per THREAD (shanten of hand +- tile from the carried base-5 signatures), effective
//              tiles become edges at the offset their rank gives (the reference's iteration order), children are interned
//              one per thread; edge slots and next-level work-list positions are reserved once per batch.
//   evaluate   levels in reverse; one thread per (state, turn) accumulates over the state's edges in the reference's order
//              with explicitly rounded f32 ops, so every tenpai / win / EV value follows the same sequence of roundings
//              as the Rust code (bit-exact); the probability vectors of a state are staged once per batch in shared memory.
//   finalize   one warp per row: candidates, comparators, rows 889..1011 written into the obs tensor
#pragma once
#include "mjx_obs.cuh"

namespace mjx {

constexpr int SP_T_MAX = 17;             // sp/mod.rs:42 MAX_TSUMOS_LEFT
constexpr int SP_SHANTEN_THRES = 3;      // calc.rs:13
constexpr int SP_MAX_TILES_LEFT = 34 * 4 - 1 - 13;  // calc.rs:14
constexpr int SP_SLOTS = 8;
constexpr int SP_B = 32;                 // states per CTA batch
constexpr int SP_THREADS = 128;
constexpr int SP_MAX_EDGES = 37;         // 34 tile kinds + 3 aka splits
constexpr u64 SP_EMPTY = ~0ull;
constexpr u32 SP_NO_CHILD = 0xFFFFFFFFu;
constexpr u32 SP_DR_NONE = 0x3FFFFu, SP_DI_NONE = 0xFFFFFFu;
constexpr int SP_VALS = 3 * SP_T_MAX;    // floats per state: tenpai[17] | win[17] | ev[17]

#ifdef MJX_HOST_EMUL
#define SP_FMUL(a, b) ((a) * (b))
#define SP_FADD(a, b) ((a) + (b))
#define SP_FDIV(a, b) ((a) / (b))
#define SP_SYNC() ((void)0)
#else
// never contracted into FMA: the reference rounds after every multiply and add
#define SP_FMUL(a, b) __fmul_rn((a), (b))
#define SP_FADD(a, b) __fadd_rn((a), (b))
#define SP_FDIV(a, b) __fdiv_rn((a), (b))
#define SP_SYNC() __syncthreads()
#endif
// block-parallel loop: every phase of a batch is `SP_PFOR` + `SP_SYNC()` (the host emulation runs one thread)
#define SP_PFOR(i, n) for (int i = B.tid; i < (n); i += B.nthr)

// sp/state.rs:10-21 of the ROOT of a row (n_extra_tsumo is always 0 without tegawari)
struct alignas(8) SpKey {
    u8 tehai[34];
    u8 wall[34];
    u8 akas;      // bits 0-2 akas_in_hand, bits 3-5 akas_in_wall
    u8 pad_[3];
};
static_assert(sizeof(SpKey) == 72, "SpKey layout");

// ---- the 64-bit state key
MJX_HD u64 sp_key_make(u32 row, u32 dr, u32 di) { return ((u64)row << 42) | ((u64)di << 18) | (u64)dr; }
MJX_HD u32 sp_key_row(u64 k) { return (u32)(k >> 42); }
MJX_HD u32 sp_key_dr(u64 k) { return (u32)k & SP_DR_NONE; }
MJX_HD u32 sp_key_di(u64 k) { return (u32)(k >> 18) & SP_DI_NONE; }
// number of draws a state is away from the root of its row. Only turns >= that depth of a state's value vectors are ever read:
// the root is evaluated at turns 0..T-1, a discard level reads its children at the same turn (calc.rs:563-637), a draw level at
// turn i reads its children at turns j+1 > i (calc.rs:486-530) -- so a state n draws down is only asked about turns >= n and
// the evaluation skips the others (their slots keep stale values nobody reads).
MJX_HD int sp_key_depth(u64 k) {
    const u32 dr = (u32)k & SP_DR_NONE;
    return (int)((dr & 63u) != 63u) + (int)(((dr >> 6) & 63u) != 63u) + (int)(((dr >> 12) & 63u) != 63u);
}
// insert tile id `t` (0..36) into a tuple of `n` ascending 6-bit fields (63 = empty, always at the top)
MJX_HD u32 sp_tuple_insert(u32 packed, int n, u32 t) {
    int p = 0;
    for (int i = 0; i < n; i++) p += ((packed >> (6 * i)) & 63u) < t;
    const u32 low = packed & ((1u << (6 * p)) - 1u);
    const u32 high = (packed >> (6 * p)) << (6 * (p + 1));
    return (low | (t << (6 * p)) | high) & ((n == 3) ? SP_DR_NONE : SP_DI_NONE);
}
// how many fields of the tuple are tile kind `t` (aka ids count as their 5)
MJX_HD int sp_tuple_count(u32 packed, int n, int t) {
    int c = 0;
    for (int i = 0; i < n; i++) {
        const int f = (int)((packed >> (6 * i)) & 63u);
        c += f != 63 && deaka(f) == t;
    }
    return c;
}
// akas byte of the state: the root's, with drawn akas moved wall -> hand and discarded akas removed
MJX_HD int sp_key_akas(int root_akas, u32 dr, u32 di) {
    int a = root_akas;
    for (int i = 0; i < 3; i++) {
        const int f = (int)((dr >> (6 * i)) & 63u);
        if (f >= T_5MR && f <= T_5SR) a = (a | (1 << (f - T_5MR))) & ~(1 << (3 + f - T_5MR));
    }
    for (int i = 0; i < 4; i++) {
        const int f = (int)((di >> (6 * i)) & 63u);
        if (f >= T_5MR && f <= T_5SR) a &= ~(1 << (f - T_5MR));
    }
    return a;
}
MJX_HD u32 sp_hash64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (u32)x;
}

// shanten signatures of a state's hand (mjx_algo.cuh HandSig), 16 bytes: w[s] = base-5 suit index (21 bits) | counter << 21
struct alignas(16) SpSigP { u32 w[4]; };
MJX_D SpSigP sp_sig_pack(const HandSig& h) {
    SpSigP p;
    p.w[0] = h.idx[0] | ((u32)h.kinds << 21); p.w[1] = h.idx[1] | ((u32)h.pairs << 21);
    p.w[2] = h.idx[2] | ((u32)h.kkinds << 21); p.w[3] = h.idx[3] | ((u32)h.kpairs << 21);
    return p;
}
MJX_D HandSig sp_sig_unpack(const SpSigP& p) {
    HandSig h;
    for (int i = 0; i < 4; i++) h.idx[i] = p.w[i] & 0x1FFFFFu;
    h.kinds = (int)(p.w[0] >> 21); h.pairs = (int)(p.w[1] >> 21); h.kkinds = (int)(p.w[2] >> 21); h.kpairs = (int)(p.w[3] >> 21);
    return h;
}

// shanten (shanten.rs:138-150 calc_all) of the hand behind `sg` with tile t added (ADD) or removed; c0 = count of t before.
// Only suit t / 9 changes: one table gather + entry 5 + len of the merge with `others` (the other three suits pre-merged);
// the chiitoi / kokushi counters are touched only when they can matter.
template <bool ADD>
MJX_D int sp_candidate_shanten(const Tables& T, const SpSigP& sg, const u8* others, int t, int c0, int len_div3) {
    const int sx = t / 9, pos = t - 9 * sx;
    const u32 w5 = sx < 3 ? c_pow5[pos] : c_pow5[pos + 2];
    const u32 idx = (sg.w[sx] & 0x1FFFFFu) + (ADD ? w5 : 0u - w5);
    const u64 r = sx < 3 ? ld_row(T.suhai, idx, SUHAI_ROWS) : ld_row(T.jihai, idx, JIHAI_ROWS);
    int v = 1 << 20;
#pragma unroll
    for (int a = 0; a <= 4; a++) {
        if (a <= len_div3) {
            const int b = len_div3 - a;
            v = min(v, (int)others[5 + a] + (int)((r >> (4 * b)) & 0xF));
            v = min(v, (int)others[a] + (int)((r >> (4 * (5 + b))) & 0xF));
        }
    }
    int sh = v - 1;
    if (sh <= 0 || len_div3 < 4) return sh;
    int kinds = (int)(sg.w[0] >> 21), pairs = (int)(sg.w[1] >> 21), kkinds = (int)(sg.w[2] >> 21), kpairs = (int)(sg.w[3] >> 21);
    const bool yao = is_yaokyuu(t);
    if (ADD) { kinds += c0 == 0; pairs += c0 == 1; if (yao) { kkinds += c0 == 0; kpairs += c0 == 1; } }
    else { kinds -= c0 == 1; pairs -= c0 == 2; if (yao) { kkinds -= c0 == 1; kpairs -= c0 == 2; } }
    sh = min(sh, 7 - pairs + max(7 - kinds, 0) - 1);
    if (sh > 0) sh = min(sh, 14 - kkinds - (kpairs > 0 ? 1 : 0) - 1);
    return sh;
}

// per observation row: sp/calc.rs:36-62 parameters + what obs_repr.rs needs afterwards
struct SpRow {
    u8 tehai_len_div3, is_menzen, prefer_riichi, calc_double_riichi, calc_haitei;
    u8 bakaze, jikaze, num_doras_in_fuuro, n_dora;
    u8 dora_ind[5];
    u8 melds[16];  // chis | pons | minkans | ankans
    u8 n_chis, n_pons, n_minkans, n_ankans;
    u8 T;          // tsumos_left = MAX_TSUMO
    u8 n_left;     // tiles in the wall at the root
    u8 avail;      // single_player_tables() is Ok
    u8 has_values; // cur_shanten <= 3
    u8 can_discard;   // effective flag handed to SPCalculator::calc (false after an accepted riichi)
    u8 cd_flag;       // cans.can_discard (what obs_repr.rs branches on)
    u8 after_riichi, last_self_tsumo;
    i8 cur_shanten;
    u8 seat;
    i32 table;
    u32 root;         // table slot of the root state, SP_NO_CHILD if none
    float fallback_ev;  // obs_repr.rs:604-616
    SpKey root_key;
};

struct SpGlobal {
    SpRow* rows;        // [row_cap]
    u64* hkey;          // [hash_cap] state key (SP_EMPTY = free); the slot index is the state's id
    SpSigP* nsig;       // [hash_cap] shanten signatures of the state's hand
    u64* einfo;         // [hash_cap] BY DENSE ID: edge_begin | n_edges << 32 | sum of required-tile counts << 40 | counts present (4 bits) << 48
    u64* dkey;          // [hash_cap] BY DENSE ID: the state's key (what the evaluation needs of hkey without the scattered read)
    float* vals;        // [hash_cap][3][SP_T_MAX] value vectors, indexed by the DENSE state id (sp_vid): level base + work-list position
    u32* sid;           // [hash_cap] level << 28 | work-list position of the state in that slot (written by whoever appends it)
    u32* echild;        // [edge_cap] child state (table slot)
    u32* evid;          // [edge_cap] dense id of the child (k_sp_densify, after the expansion): what the evaluation follows
    u16* emeta;         // [edge_cap] tile (6 bits) | count << 6 | (no-yaku flag << 15, tenpai states only)
    u32* eowner;        // [edge_cap] owning state (dense id); written for the tenpai (W0) level only (k_sp_score walks that edge range)
    float* leaf_scores; // [score_cap][4] get_score of every winning draw of the tenpai (W0) states
    u32* wl;            // [SP_SLOTS][wl_cap] work list of each level: table slots
    i32* wl_count;      // [SP_SLOTS]
    const float* p_tab;    // [SP_NTS_DIM][SP_NTS_DIM][4][SP_TRI] draw probabilities by (tiles left, sum of required counts, count, i, j)
    i32* counters;      // [1] edges, [2] overflow flag, [3] overflow events (cumulative), [4],[5] edge range of the tenpai level
    i32 row_base;       // rows [row_base, ...) of the step form this DP (row groups of mjx_env_encode_obs_host)
    i32 hash_cap, wl_cap, edge_cap, score_cap;
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

struct SpBlk { int tid, nthr, bid, nblk; };  // thread / block coordinates of a launch (host emulation: 0, 1, 0, 1)

MJX_HD bool sp_slot_is_w(int slot) { return (slot & 1) != 0; }
MJX_HD int sp_slot_shanten(int slot) { return 3 - (slot >> 1); }
MJX_D float* sp_vals(const SpGlobal& G, u32 vid, int which) { return G.vals + ((size_t)vid * 3 + which) * SP_T_MAX; }  // vid: dense id
MJX_D u32 sp_einfo_begin(u64 e) { return (u32)e; }
MJX_D int sp_einfo_n(u64 e) { return (int)((e >> 32) & 0xFF); }
MJX_D int sp_einfo_sum(u64 e) { return (int)((e >> 40) & 0xFF); }
MJX_D int sp_einfo_cmask(u64 e) { return (int)((e >> 48) & 0xF); }

MJX_D void sp_set_overflow(const SpGlobal& G) { G.counters[2] = 1; }

// Dense state ids. The table slot of a state is a hash of its key, so value vectors indexed by slot would be scattered over the
// whole table (gigabytes) and every child read of the evaluation would miss L2. The work lists already enumerate the states
// level by level, so the value vectors live at `level base + position in the level's work list`: the children of a level are
// one contiguous, just-written region (tens of MB: L2-resident) instead.
constexpr u32 SP_SID_POS = (1u << 28) - 1u;
MJX_D u32 sp_level_base(const SpGlobal& G, int level) {
    u32 b = 0;
    for (int l = 0; l < level; l++) b += (u32)min(G.wl_count[l], G.wl_cap);
    return b;
}
MJX_D u32 sp_vid(const SpGlobal& G, u32 slot) {
    const u32 s = G.sid[slot];
    return sp_level_base(G, (int)(s >> 28)) + (s & SP_SID_POS);
}
// after the expansion (all work lists final): the dense id of every edge's child
MJX_DN void sp_densify(const SpGlobal& G, const SpBlk& B) {
    if (G.counters[2]) return;  // overflow: the step's block is dropped anyway and the evaluation does not run
    u32 base[SP_SLOTS];
    for (int l = 0; l < SP_SLOTS; l++) base[l] = sp_level_base(G, l);
    const int n = min(G.counters[1], G.edge_cap), leaf_b = G.counters[4], leaf_e = G.counters[5];
    for (int e = B.bid * B.nthr + B.tid; e < n; e += B.nblk * B.nthr) {
        if (e >= leaf_b && e < leaf_e) continue;  // winning draws have no child
        const u32 c = G.echild[e];
        u32 v = SP_NO_CHILD;
        if (c != SP_NO_CHILD) { const u32 sd = G.sid[c]; v = base[sd >> 28] + (sd & SP_SID_POS); }
        G.evid[e] = v;
    }
}

// pull a state's value vectors (204 B: two 128-byte lines) towards the SM ahead of their use
MJX_D void sp_prefetch(const float* p) {
#ifndef MJX_HOST_EMUL
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
#else
    (void)p;
#endif
}
MJX_D i32 sp_atomic_add(i32* p, i32 v) {
#ifdef MJX_HOST_EMUL
    const i32 o = *p; *p += v; return o;
#else
    return atomicAdd(p, v);
#endif
}
MJX_D void sp_atomic_or(u32* p, u32 v) {
#ifdef MJX_HOST_EMUL
    *p |= v;
#else
    atomicOr(p, v);
#endif
}

// find-or-insert a state key; the returned slot is the state's id. `won` = this thread created it (and must append it to
// the next level's work list and write its signature). Lock-free: one 64-bit CAS publishes the state, the key IS the entry.
MJX_DN u32 sp_intern(const SpGlobal& G, u64 key, bool& won) {
    const u32 mask = (u32)G.hash_cap - 1;
    u32 h = sp_hash64(key) & mask;
    won = false;
    for (int probe = 0; probe < 8192; probe++, h = (h + 1) & mask) {
#ifdef MJX_HOST_EMUL
        const u64 cur = G.hkey[h];
        if (cur == key) return h;
        if (cur == SP_EMPTY) { G.hkey[h] = key; won = true; return h; }
#else
        const u64 cur = __ldcg(reinterpret_cast<const unsigned long long*>(G.hkey + h));
        if (cur == key) return h;
        if (cur == SP_EMPTY) {
            const u64 prev = atomicCAS(reinterpret_cast<unsigned long long*>(G.hkey + h), (unsigned long long)SP_EMPTY,
                                       (unsigned long long)key);
            if (prev == SP_EMPTY) { won = true; return h; }
            if (prev == key) return h;
        }
#endif
    }
    sp_set_overflow(G);  // table (nearly) full
    return SP_NO_CHILD;
}

// ---------------------------------------------------------------------------------------------- expansion
// shared-memory working set of one batch
struct SpExpandBatch {
    u64 key[SP_B];
    u32 slot[SP_B];
    SpSigP sig[SP_B];
    u8 akas[SP_B], len[SP_B];
    u8 teh[SP_B][34], wall[SP_B][34];
    u8 oth[SP_B][4][12];   // per suit s: min-plus merge of the OTHER three suits' table rows (10 entries)
    u32 cand[SP_B][2], eff[SP_B][2];
    u8 split[SP_B], ne[SP_B], sumreq[SP_B], cmask[SP_B];
    u16 cand_off[SP_B + 1], e_off[SP_B + 1];
    u16 cand_list[SP_B * 34];
    u32 edge[SP_B * SP_MAX_EDGES];  // state | tile << 8 | count << 16
    u32 win[SP_B * SP_MAX_EDGES];
    i32 n_cand, n_edge, e_base, n_win, w_base;
};

MJX_D u64 sp_mask64(const u32* m) { return (u64)m[0] | ((u64)m[1] << 32); }

// KIND 0: D level (shanten-keeping discards, sp/state.rs:98-137), 1: W level (shanten-lowering draws, state.rs:139-179),
// 2: the tenpai W level (winning draws: edges only, no children)
template <int KIND>
MJX_DN void sp_expand_batch(const SpGlobal& G, const Tables& T, SpExpandBatch& S, const SpBlk& B, int level, int first, int nb) {
    constexpr bool IS_W = KIND != 0, LEAF = KIND == 2;
    const int k = sp_slot_shanten(level);
    const u32* list = G.wl + (size_t)level * G.wl_cap;
    SP_PFOR(st, nb) {
        const u32 slot = list[first + st];
        const u64 key = G.hkey[slot];
        const SpRow& R = G.rows[sp_key_row(key)];
        S.slot[st] = slot; S.key[st] = key;
        S.sig[st] = G.nsig[slot];
        S.akas[st] = (u8)sp_key_akas(R.root_key.akas, sp_key_dr(key), sp_key_di(key));
        S.len[st] = R.tehai_len_div3;
        S.cand[st][0] = S.cand[st][1] = S.eff[st][0] = S.eff[st][1] = 0;
    }
    if (B.tid == 0) S.n_win = 0;
    SP_SYNC();
    // counts of every tile kind in the state's hand and wall: the root's, then the state's <= 7 draws / discards applied
    SP_PFOR(it, nb * 34) {
        const int st = it / 34, t = it - st * 34;
        const SpRow& R = G.rows[sp_key_row(S.key[st])];
        S.teh[st][t] = R.root_key.tehai[t]; S.wall[st][t] = R.root_key.wall[t];
    }
    // for every suit the min-plus merge of the other three suits' rows (shanten.rs:51-80 is an exact min-plus convolution,
    // so it may be associated freely): a candidate then costs ONE table gather and one partial merge instead of four and three
    SP_PFOR(it, nb * 4) {
        const int st = it >> 2, sx = it & 3;
        const HandSig h = sp_sig_unpack(S.sig[st]);
        Row10 acc;
        bool first = true;
        for (int q = 0; q < 4; q++) {
            if (q == sx) continue;
            const Row10 r = unpack_row(q < 3 ? ld_row(T.suhai, h.idx[q], SUHAI_ROWS) : ld_row(T.jihai, h.idx[q], JIHAI_ROWS));
            if (first) { acc = r; first = false; } else add_suhai_full(acc, r);
        }
        for (int q = 0; q < 10; q++) S.oth[st][sx][q] = (u8)acc.v[q];
    }
    SP_SYNC();
    // candidate tiles: still in the wall (W) / held (D)
    SP_PFOR(st, nb) {
        const u64 key = S.key[st];
        const u32 dr = sp_key_dr(key), di = sp_key_di(key);
        for (int i = 0; i < 3; i++) { const int f = (int)((dr >> (6 * i)) & 63u); if (f != 63) { S.teh[st][deaka(f)] += 1; S.wall[st][deaka(f)] -= 1; } }
        for (int i = 0; i < 4; i++) { const int f = (int)((di >> (6 * i)) & 63u); if (f != 63) S.teh[st][deaka(f)] -= 1; }
        u64 m = 0;
        for (int t = 0; t < 34; t++) if ((IS_W ? S.wall[st][t] : S.teh[st][t]) > 0) m |= 1ull << t;
        S.cand[st][0] = (u32)m; S.cand[st][1] = (u32)(m >> 32);
    }
    SP_SYNC();
    if (B.tid == 0) {
        int acc = 0;
        for (int st = 0; st < nb; st++) { S.cand_off[st] = (u16)acc; acc += mjx_popcll(sp_mask64(S.cand[st])); }
        S.cand_off[nb] = (u16)acc; S.n_cand = acc;
    }
    SP_SYNC();
    SP_PFOR(it, nb * 34) {
        const int st = it / 34, t = it - st * 34;
        const u64 m = sp_mask64(S.cand[st]);
        if ((m >> t) & 1) S.cand_list[S.cand_off[st] + mjx_popcll(m & ((1ull << t) - 1))] = (u16)((st << 6) | t);
    }
    SP_SYNC();
    // one candidate per thread: shanten of hand +- tile
    SP_PFOR(i, S.n_cand) {
        const int st = S.cand_list[i] >> 6, t = S.cand_list[i] & 63;
        const int sh = sp_candidate_shanten<IS_W>(T, S.sig[st], S.oth[st][t / 9], t, S.teh[st][t], S.len[st]);
        const bool ok = IS_W ? sh - k == -1 : sh == k;
        if (ok) sp_atomic_or(&S.eff[st][t >> 5], 1u << (t & 31));
    }
    SP_SYNC();
    // an effective 5 whose aka is still in the wall splits in two edges (sp/state.rs:160-176)
    SP_PFOR(st, nb) {
        const u64 eff = sp_mask64(S.eff[st]);
        int split = 0, sum = 0, cmask = 0;  // cmask: which edge counts 1..4 occur (the evaluation caches one probability row each)
        if (IS_W) {
            for (int s5 = 0; s5 < 3; s5++) {
                const int t5 = 4 + 9 * s5;
                if (((eff >> t5) & 1) && ((S.akas[st] >> (3 + s5)) & 1) && S.wall[st][t5] >= 2) split |= 1 << s5;
            }
            for (u64 rest = eff; rest; rest &= rest - 1) {
                const int t = mjx_ffsll(rest) - 1, cnt = S.wall[st][t];
                sum += cnt;
                const int s5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
                if (s5 >= 0 && ((S.akas[st] >> (3 + s5)) & 1)) { cmask |= 1; if (cnt >= 2) cmask |= 1 << (cnt - 2); }
                else cmask |= 1 << (cnt - 1);
            }
        }
        S.split[st] = (u8)split;
        S.cmask[st] = (u8)cmask;
        S.ne[st] = (u8)(mjx_popcll(eff) + mjx_popc((u32)split));
        S.sumreq[st] = (u8)sum;  // u8 arithmetic, as the reference's `.sum::<u8>()`
    }
    SP_SYNC();
    if (B.tid == 0) {
        int acc = 0;
        for (int st = 0; st < nb; st++) { S.e_off[st] = (u16)acc; acc += S.ne[st]; }
        S.e_off[nb] = (u16)acc;
        int eb = sp_atomic_add(&G.counters[1], acc);
        if (eb + acc > G.edge_cap) { sp_set_overflow(G); acc = 0; eb = 0; }
        S.n_edge = acc; S.e_base = eb;
    }
    SP_SYNC();
    SP_PFOR(st, nb) {
        const int ne = S.n_edge ? S.ne[st] : 0;
        const u32 vid = sp_level_base(G, level) + (u32)(first + st);  // this level's work list is final: so is the state's dense id
        G.einfo[vid] = (u64)(u32)(S.e_base + S.e_off[st]) | ((u64)ne << 32) | ((u64)S.sumreq[st] << 40) | ((u64)S.cmask[st] << 48);
        G.dkey[vid] = S.key[st];
    }
    // edge descriptors in tile order: every effective tile writes its own at the offset its rank gives
    if (S.n_edge) SP_PFOR(it, nb * 34) {
        const int st = it / 34, t = it - st * 34;
        const u64 eff = sp_mask64(S.eff[st]);
        if (!((eff >> t) & 1)) continue;
        int off = S.e_off[st] + mjx_popcll(eff & ((1ull << t) - 1));
        for (int s5 = 0; s5 < 3; s5++) off += ((S.split[st] >> s5) & 1) && 4 + 9 * s5 < t;
        const int suit5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
        if (IS_W) {
            const int count = S.wall[st][t];
            if (suit5 >= 0 && ((S.akas[st] >> (3 + suit5)) & 1)) {
                if (count >= 2) { S.edge[off] = (u32)st | ((u32)t << 8) | ((u32)(count - 1) << 16); off++; }
                S.edge[off] = (u32)st | ((u32)(T_5MR + suit5) << 8) | (1u << 16);
            } else {
                S.edge[off] = (u32)st | ((u32)t << 8) | ((u32)count << 16);
            }
        } else {
            // a discarded 5 is the aka only when it is the last 5 in hand (state.rs:127-132)
            int tile = t;
            if (suit5 >= 0 && ((S.akas[st] >> suit5) & 1) && S.teh[st][t] == 1) tile = T_5MR + suit5;
            S.edge[off] = (u32)st | ((u32)tile << 8);
        }
    }
    SP_SYNC();
    // one edge per thread: child key and signatures (both incremental), interning
    SP_PFOR(e, S.n_edge) {
        const u32 d = S.edge[e];
        const int st = (int)(d & 0xFF), tile = (int)((d >> 8) & 0xFF), cnt = (int)(d >> 16);
        const int ge = S.e_base + e;
        G.emeta[ge] = (u16)(tile | (cnt << 6));
        if (LEAF) { G.eowner[ge] = sp_level_base(G, level) + (u32)(first + st); continue; }
        const u64 key = S.key[st];
        const int t = deaka(tile);
        const u64 ckey = IS_W ? sp_key_make(sp_key_row(key), sp_tuple_insert(sp_key_dr(key), 3, (u32)tile), sp_key_di(key))
                              : sp_key_make(sp_key_row(key), sp_key_dr(key), sp_tuple_insert(sp_key_di(key), 4, (u32)tile));
        bool won;
        const u32 child = sp_intern(G, ckey, won);
        G.echild[ge] = child;
        if (won) {
            G.nsig[child] = sp_sig_pack(sig_variant(sp_sig_unpack(S.sig[st]), t, IS_W ? +1 : -1, S.teh[st][t]));
            S.win[sp_atomic_add(&S.n_win, 1)] = child;
        }
    }
    SP_SYNC();
    if (!LEAF) {
        if (B.tid == 0) {
            int wb = sp_atomic_add(&G.wl_count[level + 1], S.n_win);
            if (wb + S.n_win > G.wl_cap) { sp_set_overflow(G); S.n_win = 0; wb = 0; }
            S.w_base = wb;
        }
        SP_SYNC();
        u32* next = G.wl + (size_t)(level + 1) * G.wl_cap;
        SP_PFOR(i, S.n_win) {
            next[S.w_base + i] = S.win[i];
            G.sid[S.win[i]] = ((u32)(level + 1) << 28) | (u32)(S.w_base + i);
        }
        SP_SYNC();
    }
}

template <int KIND>
MJX_DN void sp_expand_level(const SpGlobal& G, const Tables& T, SpExpandBatch& S, const SpBlk& B, int level) {
    const int n = min(G.wl_count[level], G.wl_cap);
    for (int first = B.bid * SP_B; first < n; first += B.nblk * SP_B) sp_expand_batch<KIND>(G, T, S, B, level, first, min(SP_B, n - first));
}

// ---------------------------------------------------------------------------------------------- scoring of winning draws
// calc.rs:640-758 for one winning draw; executed by one thread. Returns false when there is no yaku.
MJX_DN bool sp_get_score(const Tables& T, const SpRow& P, const u8* tehai13, const u8* wall_in, int akas, int win_tile, float* scores) {
    u8 th[34];
    for (int i = 0; i < 34; i++) th[i] = tehai13[i];
    const int wid = deaka(win_tile);
    th[wid] += 1;
    const int akas_in_hand = (akas & 7) | (is_aka(win_tile) ? (1 << (win_tile - T_5MR)) : 0);
    u8 wall[34];
    for (int i = 0; i < 34; i++) wall[i] = wall_in[i];
    wall[wid] -= 1;
    AgariQuery q;
    q.tehai = th;
    q.chis = P.melds; q.pons = P.melds + 4; q.minkans = P.melds + 8; q.ankans = P.melds + 12;
    q.n_chis = P.n_chis; q.n_pons = P.n_pons; q.n_minkans = P.n_minkans; q.n_ankans = P.n_ankans;
    q.bakaze = P.bakaze; q.jikaze = P.jikaze; q.winning_tile = wid; q.is_ron = false; q.is_menzen = P.is_menzen != 0;
    const bool is_oya = P.jikaze == T_E;
    const int additional = P.is_menzen ? (P.prefer_riichi ? 2 : 1) : 0;
    int doras = mjx_popc((u32)akas_in_hand) + P.num_doras_in_fuuro;
    for (int i = 0; i < P.n_dora; i++) doras += th[tile_next(P.dora_ind[i])];
    Agari a = agari_with(T, q, additional, doras & 0xFF);
    if (a.kind == 0) return false;
    if (a.kind == 2) {
        float v = (float)tsumo_total(point_yakuman(is_oya, a.yakuman), is_oya);
        for (int i = 0; i < 4; i++) scores[i] = v;
        return true;
    }
    const int fu = a.fu, han = a.han;
    const bool assume_riichi = P.is_menzen && P.prefer_riichi;
    // tsumo totals for han .. han+15 once (the reference recomputes Point::calc inside the double loops)
    const int n_extra = !assume_riichi ? 4 : (P.n_dora == 1 ? 8 : 16);
    float pts[16];
    for (int h = 0; h < n_extra; h++) { bool ok; pts[h] = (float)tsumo_total(point_calc(is_oya, fu, han + h, &ok), is_oya); }
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
                scores[i] = SP_FADD(scores[i], SP_FMUL(pts[i + j], up[j]));
            }
    } else if (assume_riichi && P.n_dora > 1) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 13; j++) {
                float p = c_uradora_prob[P.n_dora - 1][j];
                if (p == 0.f) continue;
                scores[i] = SP_FADD(scores[i], SP_FMUL(pts[i + j], p));
            }
    } else {
        for (int i = 0; i < 4; i++) scores[i] = pts[i];
    }
    return true;
}

// score one winning draw of a tenpai state (calc.rs:478-479 get_score), ONE THREAD PER DRAW; the per-turn accumulation
// happens in the evaluation of the tenpai level.
MJX_DN void sp_score_edge(const SpGlobal& G, const Tables& T, int e) {
    const u32 node = G.eowner[e];
    const u64 key = G.dkey[node];
    const SpRow& P = G.rows[sp_key_row(key)];
    const u32 dr = sp_key_dr(key), di = sp_key_di(key);
    u8 th[34], wall[34];
    for (int t = 0; t < 34; t++) { th[t] = P.root_key.tehai[t]; wall[t] = P.root_key.wall[t]; }
    for (int i = 0; i < 3; i++) { const int f = (int)((dr >> (6 * i)) & 63u); if (f != 63) { th[deaka(f)] += 1; wall[deaka(f)] -= 1; } }
    for (int i = 0; i < 4; i++) { const int f = (int)((di >> (6 * i)) & 63u); if (f != 63) th[deaka(f)] -= 1; }
    float sc[4];
    const u16 m = G.emeta[e];
    const bool ok = sp_get_score(T, P, th, wall, sp_key_akas(P.root_key.akas, dr, di), m & 63, sc);
    const int le = e - G.counters[4];
    if (le >= G.score_cap) G.counters[2] = 1;  // score arena too small for this step: reported as an overflow
    if (!ok) G.emeta[e] = (u16)(m | 0x8000);
    else if (le >= 0 && le < G.score_cap) {
        float* o = G.leaf_scores + (size_t)le * 4;
        o[0] = sc[0]; o[1] = sc[1]; o[2] = sc[2]; o[3] = sc[3];
    }
}

// ---------------------------------------------------------------------------------------------- evaluation
struct SpEvalDBatch {
    u32 vid[SP_B], ebeg[SP_B];  // vid: dense id (where the state's values go)
    u8 ne[SP_B], T[SP_B], d[SP_B];  // d: first turn anybody reads (sp_key_depth)
    u16 off[SP_B + 1];
    i32 n_items;
};
// W levels. The probability that the draw of turn j brings an effective tile of count c when turn i is the current one
// (calc.rs:486-497 `tsumo_probs[j] * not_tsumo_probs[j] / not_tsumo_probs[i]`, with the tables of calc.rs:136-167) depends on
// (tiles left at the root, sum of required counts, c, i, j) only — 124 x 124 x 4 x 153 values. They are tabulated ONCE per
// process with the reference's own operation sequence (k_sp_tables; 37.6 MB, L2-resident in its hot part), so the inner loop of
// the evaluation performs the reference's rounded multiply-adds with a table load in place of two divisions and a multiply.
constexpr int SP_NTS_DIM = SP_MAX_TILES_LEFT + 2;  // n_left, sum_required in 0..123
constexpr int SP_TRI = SP_T_MAX * (SP_T_MAX + 1) / 2;  // (i, j) pairs, i <= j < 17: row i (j contiguous) starts at sp_tri_row(i)
MJX_HD int sp_tri_row(int i) { return SP_T_MAX * i - i * (i - 1) / 2; }
MJX_D size_t sp_ptab_index(int n_left, int i0) { return ((size_t)n_left * SP_NTS_DIM + (size_t)i0) * 4 * SP_TRI; }
// one (n_left, i0) block of the table: [c][pair]
MJX_D void sp_fill_ptab_block(float* blk, int n_left, int i0) {
    float nts[SP_T_MAX];  // not_tsumo_prob_table[i0]: row[0] = 1, row[j+1] = row[j] * (n_left - i0 - j) / (n_left - j)  (calc.rs:158-165)
    float v = 1.f;
    for (int j = 0; j < SP_T_MAX; j++) {
        const bool ok = i0 <= n_left && j <= n_left - i0;
        nts[j] = ok ? v : 0.f;
        if (ok && j < n_left - i0) v = SP_FDIV(SP_FMUL(v, (float)(n_left - i0 - j)), (float)(n_left - j));
    }
    for (int c = 0; c < 4; c++)
        for (int j = 0; j < SP_T_MAX; j++) {
            // tsumo_prob_table[c][j] = (c + 1) / (n_left - j)  (calc.rs:136-146)
            const float tpj = (n_left - j > 0 && nts[j] != 0.f) ? SP_FMUL(SP_FDIV((float)(c + 1), (float)(n_left - j)), nts[j]) : 0.f;
            for (int i = 0; i <= j; i++) blk[c * SP_TRI + sp_tri_row(i) + (j - i)] = nts[i] != 0.f ? SP_FDIV(tpj, nts[i]) : 0.f;
        }
}

constexpr int SP_WB = 7;  // W states per WARP mini-batch: 7 x ceil(17 / 2) = 63 (state, turn-pair) items = two rounds of 32 lanes (16 per warp measured slower)
struct SpEvalWBatch {      // one per warp: the W evaluation needs no CTA-wide barrier
    u32 vid[SP_WB], ebeg[SP_WB], pbase[SP_WB];  // vid: dense id (where the state's values go)
    u8 ne[SP_WB], T[SP_WB], d[SP_WB], flags[SP_WB], jend[SP_WB];  // d: first turn anybody reads (sp_key_depth); flags: 1 assume_riichi, 2 double riichi, 4 haitei; jend: first j with not_tsumo[j] == 0
    u16 aoff[SP_WB + 1];
    i32 n_a;
};
#ifdef MJX_HOST_EMUL
#define SP_WSYNC() ((void)0)
#else
#define SP_WSYNC() __syncwarp()
#endif
template <typename Tb>
MJX_D int sp_find_state(const u16* off, int nb, int item) {
    int lo = 0, hi = nb - 1;  // last state whose offset is <= item
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)off[mid] <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// calc.rs:447-561 draw_without_tegawari_slow. LEAF: the tenpai level (scores of the winning draws); otherwise shanten k >= 1.
template <bool LEAF>
MJX_DN void sp_eval_w_batch(const SpGlobal& G, SpEvalWBatch& S, const SpBlk& B, int level, int first, int nb) {
    const int k = LEAF ? 0 : sp_slot_shanten(level);
    const u32 vbase = sp_level_base(G, level) + (u32)first;  // the states of the batch are consecutive dense ids
    SP_PFOR(st, nb) {
        const u64 key = G.dkey[vbase + st];
        const SpRow& R = G.rows[sp_key_row(key)];
        const u64 ei = G.einfo[vbase + st];
        const int Tn = R.T, n_left = R.n_left, i0 = sp_einfo_sum(ei);
        const bool row_ok = i0 <= n_left && i0 <= SP_MAX_TILES_LEFT;
        const int lim = row_ok ? min(Tn - 1, n_left - i0) : -1;
        S.vid[st] = vbase + (u32)st; S.ebeg[st] = sp_einfo_begin(ei); S.ne[st] = (u8)sp_einfo_n(ei);
        S.T[st] = (u8)Tn; S.d[st] = (u8)min(sp_key_depth(key), Tn);
#ifdef MJX_HOST_EMUL  // the emulated parity tests poison the turns nobody may read
        for (int i = 0; i < S.d[st]; i++) { float* o = G.vals + (size_t)S.vid[st] * SP_VALS; o[i] = o[SP_T_MAX + i] = o[2 * SP_T_MAX + i] = __builtin_nanf(""); }
#endif
        const bool ar = R.is_menzen && R.prefer_riichi;
        S.flags[st] = (u8)((ar ? 1 : 0) | (R.calc_double_riichi ? 2 : 0) | (R.calc_haitei ? 4 : 0));
        S.jend[st] = (u8)max(0, min(Tn, lim + 1));
        S.pbase[st] = row_ok ? (u32)sp_ptab_index(n_left, i0) : 0u;
    }
    SP_WSYNC();
    if (B.tid == 0) {
        int aa = 0;
        for (int st = 0; st < nb; st++) { S.aoff[st] = (u16)aa; aa += (S.T[st] - S.d[st] + 1) / 2; }
        S.aoff[nb] = (u16)aa; S.n_a = aa;
    }
    SP_WSYNC();
    // accumulation: a thread takes turns d+p and T-1-p of a state (T-d+1 draw turns together: balanced), edges in the reference's
    // order, j ascending, every multiply and add rounded as the reference rounds it
    SP_PFOR(item, S.n_a) {
        const int st = sp_find_state<int>(S.aoff, nb, item), p = item - S.aoff[st];
        const int Tn = S.T[st], ne = S.ne[st], jend = S.jend[st];
        const u32 eb = S.ebeg[st];
        const float* Pb = G.p_tab + S.pbase[st];
        // row i of the probability triangle is contiguous in j: Pc[row(i) - i + j]
        const bool assume_riichi = (S.flags[st] & 1) != 0, dbl = (S.flags[st] & 2) != 0, haitei = (S.flags[st] & 4) != 0;
        const int i0 = S.d[st] + p, i1 = Tn - 1 - p;
        const bool two = i1 > i0;
        float t0 = 0.f, w0 = 0.f, v0 = 0.f, t1 = 0.f, w1 = 0.f, v1 = 0.f;
        // the next edge's descriptor is fetched while the current one is accumulated (dependent chain meta/child -> values)
        u16 meta_n = ne ? G.emeta[eb] : (u16)0;
        u32 child_n = (!LEAF && ne) ? G.evid[eb] : 0u;
        for (int e = 0; e < ne; e++) {
            const u16 meta = meta_n;
            const u32 child = child_n;
            if (e + 1 < ne) { meta_n = G.emeta[eb + e + 1]; if (!LEAF) child_n = G.evid[eb + e + 1]; }
            if (!LEAF && e + 1 < ne && child_n != SP_NO_CHILD) sp_prefetch(G.vals + (size_t)child_n * SP_VALS);
            const float* Pc = Pb + (((meta >> 6) & 7) - 1) * SP_TRI;
            const float* P0 = Pc + sp_tri_row(i0) - i0;
            const float* P1 = Pc + sp_tri_row(i1 > 0 ? i1 : 0) - i1;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const float* cv = nullptr;
            if (LEAF) {
                if (meta & 0x8000) continue;  // no yaku
                const int le = (int)(eb + e) - G.counters[4];
                if (le < 0 || le >= G.score_cap) continue;
                const float* sc = G.leaf_scores + (size_t)le * 4;
                s0 = sc[0]; s1 = sc[1]; s2 = sc[2]; s3 = sc[3];
            } else {
                if (child == SP_NO_CHILD) continue;  // only after an overflow
                cv = G.vals + (size_t)child * SP_VALS;
            }
            // turn i0: the reference stops a turn whose not_tsumo_probs[i] is zero and a draw turn whose not_tsumo_probs[j] is zero
            // (the row is monotone: both are `>= jend`)
            if (i0 < jend)
                for (int j = i0; j < jend; j++) {
                    const float prob = P0[j];
                    if (LEAF) {
                        const int han_plus = (assume_riichi && dbl && i0 == 0) + (assume_riichi && j == i0) + (haitei && j == Tn - 1);
                        const float sv = han_plus == 0 ? s0 : (han_plus == 1 ? s1 : (han_plus == 2 ? s2 : s3));
                        w0 = SP_FADD(w0, prob);
                        v0 = SP_FADD(v0, SP_FMUL(prob, sv));
                    } else {
                        if (k == 1) t0 = SP_FADD(t0, prob);
                        if (j < Tn - 1) {
                            if (k > 1) t0 = SP_FADD(t0, SP_FMUL(prob, cv[j + 1]));
                            w0 = SP_FADD(w0, SP_FMUL(prob, cv[SP_T_MAX + j + 1]));
                            v0 = SP_FADD(v0, SP_FMUL(prob, cv[2 * SP_T_MAX + j + 1]));
                        }
                    }
                }
            if (two && i1 < jend)
                for (int j = i1; j < jend; j++) {
                    const float prob = P1[j];
                    if (LEAF) {
                        const int han_plus = (assume_riichi && j == i1) + (haitei && j == Tn - 1);  // i1 > 0
                        const float sv = han_plus == 0 ? s0 : (han_plus == 1 ? s1 : s2);
                        w1 = SP_FADD(w1, prob);
                        v1 = SP_FADD(v1, SP_FMUL(prob, sv));
                    } else {
                        if (k == 1) t1 = SP_FADD(t1, prob);
                        if (j < Tn - 1) {
                            if (k > 1) t1 = SP_FADD(t1, SP_FMUL(prob, cv[j + 1]));
                            w1 = SP_FADD(w1, SP_FMUL(prob, cv[SP_T_MAX + j + 1]));
                            v1 = SP_FADD(v1, SP_FMUL(prob, cv[2 * SP_T_MAX + j + 1]));
                        }
                    }
                }
        }
        float* o = G.vals + (size_t)S.vid[st] * SP_VALS;
        o[i0] = t0; o[SP_T_MAX + i0] = w0; o[2 * SP_T_MAX + i0] = v0;
        if (two) { o[i1] = t1; o[SP_T_MAX + i1] = w1; o[2 * SP_T_MAX + i1] = v1; }
    }
    SP_WSYNC();
}

// calc.rs:563-637 discard_slow: per turn the child with the largest (truncated) EV, ties by discard priority
MJX_DN void sp_eval_d_batch(const SpGlobal& G, SpEvalDBatch& S, const SpBlk& B, int level, int first, int nb) {
    const u32 vbase = sp_level_base(G, level) + (u32)first;
    SP_PFOR(st, nb) {
        const u64 ei = G.einfo[vbase + st];
        S.vid[st] = vbase + (u32)st; S.ebeg[st] = sp_einfo_begin(ei); S.ne[st] = (u8)sp_einfo_n(ei);
        const u64 key = G.dkey[vbase + st];
        S.T[st] = G.rows[sp_key_row(key)].T; S.d[st] = (u8)min(sp_key_depth(key), (int)S.T[st]);
#ifdef MJX_HOST_EMUL
        for (int i = 0; i < S.d[st]; i++) { float* o = G.vals + (size_t)S.vid[st] * SP_VALS; o[i] = o[SP_T_MAX + i] = o[2 * SP_T_MAX + i] = __builtin_nanf(""); }
#endif
    }
    SP_SYNC();
    if (B.tid == 0) {
        int acc = 0;
        for (int st = 0; st < nb; st++) { S.off[st] = (u16)acc; acc += S.T[st] - S.d[st]; }
        S.off[nb] = (u16)acc; S.n_items = acc;
    }
    SP_SYNC();
    SP_PFOR(item, S.n_items) {
        const int st = sp_find_state<int>(S.off, nb, item), i = S.d[st] + item - S.off[st];
        const int ne = S.ne[st];
        const u32 eb = S.ebeg[st];
        const float FMIN = -3.40282347e+38f;
        float bt = FMIN, bw = FMIN, bv = FMIN;
        int best_tile = T_UNK;
        i32 best_value = (i32)0x80000000;
        u32 child_n = ne ? G.evid[eb] : SP_NO_CHILD;
        for (int e = 0; e < ne; e++) {
            const u32 child = child_n;
            if (e + 1 < ne) { child_n = G.evid[eb + e + 1]; if (child_n != SP_NO_CHILD && i == S.d[st]) sp_prefetch(G.vals + (size_t)child_n * SP_VALS); }
            if (child == SP_NO_CHILD) continue;
            const int tile = G.emeta[eb + e] & 63;
            const float* cv = G.vals + (size_t)child * SP_VALS;
            const float v = cv[2 * SP_T_MAX + i];
            const i32 value = (i32)v;  // finite and < 2^31 here; Rust `as i32` truncates the same way
            if (value > best_value || (value == best_value && cmp_discard_priority(tile, best_tile) > 0)) {
                bt = cv[i]; bw = cv[SP_T_MAX + i]; bv = v;
                best_value = value; best_tile = tile;
            }
        }
        float* o = G.vals + (size_t)S.vid[st] * SP_VALS;
        o[i] = bt; o[SP_T_MAX + i] = bw; o[2 * SP_T_MAX + i] = bv;
    }
    SP_SYNC();
}

MJX_DN void sp_eval_d_level(const SpGlobal& G, SpEvalDBatch& S, const SpBlk& B, int level) {
    if (G.counters[2]) return;  // overflow: the dense ids were not assigned, the step's block is dropped
    const int n = min(G.wl_count[level], G.wl_cap);
    for (int first = B.bid * SP_B; first < n; first += B.nblk * SP_B) sp_eval_d_batch(G, S, B, level, first, min(SP_B, n - first));
}
// every WARP works through its own mini-batches (S = the warp's staging area): no CTA-wide barrier, a slow state only holds its warp
template <bool LEAF>
MJX_DN void sp_eval_w_level(const SpGlobal& G, SpEvalWBatch* S, const SpBlk& B, int level) {
    if (G.counters[2]) return;
    const int n = min(G.wl_count[level], G.wl_cap);
    const int lanes = B.nthr >= 32 ? 32 : B.nthr, wpc = B.nthr / lanes, warp = B.tid / lanes;
    SpBlk W; W.tid = B.tid - warp * lanes; W.nthr = lanes; W.bid = B.bid * wpc + warp; W.nblk = B.nblk * wpc;
    for (int first = W.bid * SP_WB; first < n; first += W.nblk * SP_WB)
        sp_eval_w_batch<LEAF>(G, S[warp], W, level, first, min(SP_WB, n - first));
}

// release the table slots this DP used (instead of a full-table memset per step); after an overflow states may exist that
// are in no work list, so the whole table is cleared.
MJX_DN void sp_release(const SpGlobal& G, const SpBlk& B) {
    const int gtid = B.bid * B.nthr + B.tid, gn = B.nblk * B.nthr;
    if (G.counters[2]) {
        for (int i = gtid; i < G.hash_cap; i += gn) G.hkey[i] = SP_EMPTY;
        return;
    }
    for (int level = 0; level < SP_SLOTS; level++) {
        const int n = min(G.wl_count[level], G.wl_cap);
        const u32* list = G.wl + (size_t)level * G.wl_cap;
        for (int i = gtid; i < n; i += gn) G.hkey[list[i]] = SP_EMPTY;
    }
}

// ---------------------------------------------------------------------------------------------- rows: init / finalize
struct SpCtx {  // warp-per-row stages
    SpGlobal G;
    Tables T;
    u8* df;     // 34 bytes of per-warp scratch (dora factors)
    int lane;
};

// per-candidate summary used by the obs rows
struct SpCand {
    int tile;           // as sp/candidate.rs (may be an aka id)
    int node;           // W-state (table slot) whose values are the candidate's, or -1 (simple mode)
    u32 vid;            // its dense id (where its values are)
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
MJX_DN void sp_required(const SpCtx& s, const Ctx& c, int len, const u8* tehai, const u8* wall, u64* set, int* num) {
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
MJX_D int real_time_shanten(const Tables& T, const TableState* S, int p) {
    const SeatPrivate& P = S->priv[p];
    if (!(P.cans & CAN_DISCARD)) return P.shanten;
    if (P.shanten > 0) return (P.flags & PF_HAS_NEXT_SHANTEN_DISCARD) ? P.shanten - 1 : P.shanten;
    if (P.last_self_tsumo != T_NONE) return ((P.waits >> deaka(P.last_self_tsumo)) & 1) ? -1 : 0;
    return shanten_all(T, P.tehai, P.tehai_len_div3);
}

// ---- stage: init (one warp per observation row). agent_helper.rs:509-593 up to the SPCalculator::calc call.
MJX_DN void sp_stage_init(SpCtx& s, const TableState* S, int row, int table, int seat) {
    const int p = seat;
    const SeatPrivate& PV = S->priv[p];
    const SeatPublic& PU = S->pub[p];
    const u16 cans = PV.cans;
    // dora factors of this record (for doras in melds and the fallback agari value)
    MJX_FOR_TILES(s, t) {
        int f = 0;
        for (int q = 0; q < S->n_dora; q++) f += tile_next(S->wall[60 - q]) == t;
        s.df[t] = (u8)f;
    }
    MJX_END_TILES(s);
    const u8* df = s.df;
    Ctx c;
    c.S = const_cast<TableState*>(S); c.W = nullptr; c.T = s.T; c.lane = s.lane; c.df = df;

    bool can_discard = (cans & CAN_DISCARD) != 0;
    const int cur_shanten = real_time_shanten(s.T, S, p);
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
    float fallback = 0.f;
    if (!avail && (cans & CAN_AGARI)) {
        // obs_repr.rs:604-616: agari_points(cans.can_ron_agari, &[]).tsumo_total(is_oya), 0 on Err
        bool ok;
        Point pt = agari_points(c, p, (cans & CAN_RON_AGARI) != 0, 0, &ok);
        if (ok) fallback = (float)tsumo_total(pt, p == S->oya);
    }
    if (MJX_IS_L0(s)) {
        SpRow& R = s.G.rows[row];
        R.avail = avail ? 1 : 0;
        R.fallback_ev = fallback;
        R.table = table; R.seat = (u8)seat;
        R.root = SP_NO_CHILD;
        R.cd_flag = (cans & CAN_DISCARD) ? 1 : 0;
        R.cur_shanten = (i8)cur_shanten;
        R.has_values = cur_shanten <= SP_SHANTEN_THRES ? 1 : 0;
        R.last_self_tsumo = PV.last_self_tsumo;
        if (avail) {
            R.tehai_len_div3 = PV.tehai_len_div3;
            R.is_menzen = (PV.flags & PF_IS_MENZEN) ? 1 : 0;
            R.prefer_riichi = S->scores[p] >= 1000 ? 1 : 0;
            R.calc_double_riichi = (can_discard && (PV.flags & PF_CAN_W_RIICHI)) ? 1 : 0;
            R.calc_haitei = calc_haitei ? 1 : 0;
            R.bakaze = (u8)(T_E + S->kyoku / 4);
            R.jikaze = (u8)(T_E + ((p + 4 - S->oya) & 3));
            R.n_dora = S->n_dora;
            for (int i = 0; i < 5; i++) R.dora_ind[i] = i < S->n_dora ? S->wall[60 - i] : 0;
            for (int i = 0; i < 4; i++) {
                R.melds[i] = PV.chis[i]; R.melds[4 + i] = PV.pons[i]; R.melds[8 + i] = PV.minkans[i]; R.melds[12 + i] = PV.ankans[i];
            }
            R.n_chis = PV.n_chis; R.n_pons = PV.n_pons; R.n_minkans = PV.n_minkans; R.n_ankans = PV.n_ankans;
            int nf = 0;  // agent_helper.rs:533-545
            if (!(R.is_menzen && PU.n_ankan == 0)) {
                for (int f = 0; f < PU.n_fuuro; f++)
                    for (int j = 0; j < 4; j++) { int t = PU.fuuro[f][j]; if (t != T_NONE) nf += df[deaka(t)] + (is_aka(t) ? 1 : 0); }
                for (int j = 0; j < PU.n_ankan; j++) { int t = PU.ankan[j]; nf += 4 * df[t] + ((t == T_5M || t == T_5P || t == T_5S) ? 1 : 0); }
            }
            R.num_doras_in_fuuro = (u8)nf;
            R.T = (u8)tsumos_left;
            const bool after_riichi = can_discard && ((S->riichi_accepted >> p) & 1);
            R.after_riichi = after_riichi ? 1 : 0;
            R.can_discard = (can_discard && !after_riichi) ? 1 : 0;
            SpKey root;
            int akas_hand = PV.akas_in_hand;
            for (int t = 0; t < 34; t++) root.tehai[t] = PV.tehai[t];
            if (after_riichi) {
                int lt = PV.last_self_tsumo;
                root.tehai[deaka(lt)] -= 1;
                if (is_aka(lt)) akas_hand &= ~(1 << (lt - T_5MR));
            }
            int n_left = 0;
            for (int t = 0; t < 34; t++) {
                int seen = S->public_seen[t] + PV.tehai[t];  // tiles_seen is NOT adjusted for the riichi discard
                root.wall[t] = (u8)(4 - seen);
                n_left += 4 - seen;
            }
            R.n_left = (u8)n_left;
            const int akas_seen = S->akas_public | PV.akas_in_hand;
            root.akas = (u8)((akas_hand & 7) | (((~akas_seen) & 7) << 3));
            root.pad_[0] = root.pad_[1] = root.pad_[2] = 0;
            R.root_key = root;
            if (R.has_values) {
                // the root state: no draws, no discards; unique per row, so the insertion always creates it
                const int level = 2 * (3 - cur_shanten) + (R.can_discard ? 0 : 1);
                bool won;
                const u32 slot = sp_intern(s.G, sp_key_make((u32)row, SP_DR_NONE, SP_DI_NONE), won);
                if (slot != SP_NO_CHILD) {
                    s.G.nsig[slot] = sp_sig_pack(hand_sig(root.tehai));
                    const int pos = sp_atomic_add(&s.G.wl_count[level], 1);
                    if (pos < s.G.wl_cap) { s.G.wl[(size_t)level * s.G.wl_cap + pos] = slot; s.G.sid[slot] = ((u32)level << 28) | (u32)pos; R.root = slot; }
                    else sp_set_overflow(s.G);
                }
            }
        }
    }
    MJX_SYNCWARP();
}

// ---- stage: finalize (one warp per row): candidates -> obs rows 889..1011 (obs_repr.rs:561-617, 644-692)
// `obs_row` points at this observation's [1012][34] block in global memory (rows 889.. are zero on entry).
MJX_DN void sp_stage_finalize(SpCtx& s, int row, float* obs_row) {
    const SpRow& R = s.G.rows[row];
    Ctx c;
    c.S = nullptr; c.W = nullptr; c.T = s.T; c.lane = s.lane; c.df = nullptr;
#ifdef MJX_HOST_EMUL
#define SPO_FILL(r, v) do { for (int c_ = 0; c_ < 34; c_++) obs_row[(r) * 34 + c_] = (v); } while (0)
#define SPO_ASSIGN(r, col, v) do { obs_row[(r) * 34 + (col)] = (v); } while (0)
#else
#define SPO_FILL(r, v) do { obs_row[(r) * 34 + s.lane] = (v); if (s.lane < 2) obs_row[(r) * 34 + 32 + s.lane] = (v); } while (0)
#define SPO_ASSIGN(r, col, v) do { if (s.lane == 0) obs_row[(r) * 34 + (col)] = (v); } while (0)
#endif
    if (!R.avail) {
        const float v = R.fallback_ev;
        SPO_FILL(889, fminf(fmaxf(v, 0.f), 100000.f) / 100000.f);
        SPO_FILL(890, fminf(fmaxf(v, 0.f), 30000.f) / 30000.f);
        return;
    }
    if (s.G.counters[2]) return;  // arena overflow this step: leave the block zero (counted, never silent)
    const int cur_shanten = R.cur_shanten;
    const bool has_values = R.has_values != 0, can_discard = R.can_discard != 0;
    const int T = R.T, len = R.tehai_len_div3;
    SpCand cands[14];
    int n_cands = 0;
    if (!has_values) {
        // calc.rs:281-314 analyze_*_simple
        if (can_discard) {
            const HandSig base = hand_sig(R.root_key.tehai);
#ifndef MJX_HOST_EMUL
            // Device form of the loop below (same results): the <= 14 discards are independent, so lane i first takes discard i
            // (hand signature and shanten after it), then the (discard, drawn tile) pairs are spread over the lanes 32 at a time --
            // 15 rounds of table gathers instead of 14 x (1 + 1 + 2) dependent ones, which were the kernel's long pole.
            {
                u64 held = 0;
                for (int t = 0; t < 34; t++) if (R.root_key.tehai[t]) held |= 1ull << t;
                const int nk = min(mjx_popcll(held), 14);
                // lane i: its discard
                int my_t = 0, my_after = 0;
                HandSig my_sig = base;
                if (s.lane < nk) {
                    u64 m = held;
                    for (int q = 0; q < s.lane; q++) m &= m - 1;
                    my_t = mjx_ffsll(m) - 1;
                    my_sig = sig_variant(base, my_t, -1, R.root_key.tehai[my_t]);
                    my_after = shanten_all_sig(s.T, my_sig, len);
                }
                u64 my_req = 0;
                const int n_items = nk * 34;
                for (int b0 = 0; b0 < n_items; b0 += 32) {
                    const int item = b0 + s.lane;
                    const int ci = min(item / 34, nk - 1), u = item - (item / 34) * 34;
                    HandSig g;
#pragma unroll
                    for (int q = 0; q < 4; q++) g.idx[q] = __shfl_sync(0xffffffffu, my_sig.idx[q], ci);
                    g.kinds = __shfl_sync(0xffffffffu, my_sig.kinds, ci); g.pairs = __shfl_sync(0xffffffffu, my_sig.pairs, ci);
                    g.kkinds = __shfl_sync(0xffffffffu, my_sig.kkinds, ci); g.kpairs = __shfl_sync(0xffffffffu, my_sig.kpairs, ci);
                    const int ct = __shfl_sync(0xffffffffu, my_t, ci), cafter = __shfl_sync(0xffffffffu, my_after, ci);
                    bool ok = false;
                    if (item < n_items && R.root_key.wall[u] > 0) {
                        const int cnt = (int)R.root_key.tehai[u] - (u == ct ? 1 : 0);  // copies of u held after the discard
                        ok = shanten_all_sig(s.T, sig_variant(g, u, +1, cnt), len) < cafter;
                    }
                    const unsigned bits = __ballot_sync(0xffffffffu, ok);  // bit l <-> item b0 + l
                    // lane i keeps the part of this round that belongs to its discard: items [34 i, 34 i + 34)
                    const int lo = max(34 * s.lane, b0), hi = min(34 * s.lane + 34, b0 + 32);
                    if (s.lane < nk && lo < hi) {
                        const unsigned part = (bits >> (lo - b0)) & (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u));
                        my_req |= (u64)part << (lo - 34 * s.lane);
                    }
                }
                int my_num = 0;
                for (u64 rest = my_req; rest; rest &= rest - 1) my_num += R.root_key.wall[mjx_ffsll(rest) - 1];
                for (int i = 0; i < nk; i++) {
                    SpCand& cd = cands[n_cands++];
                    const int t = __shfl_sync(0xffffffffu, my_t, i);
                    const int after = __shfl_sync(0xffffffffu, my_after, i);
                    const int k5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
                    cd.tile = (k5 >= 0 && ((R.root_key.akas >> k5) & 1) && R.root_key.tehai[t] == 1) ? T_5MR + k5 : t;
                    cd.node = -1;
                    cd.shanten_down = after - cur_shanten == 1;
                    cd.t0 = cd.w0 = cd.e0 = 0.f;
                    const u32 rlo = __shfl_sync(0xffffffffu, (u32)my_req, i), rhi = __shfl_sync(0xffffffffu, (u32)(my_req >> 32), i);
                    cd.required = (u64)rlo | ((u64)rhi << 32);
                    cd.num_required = __shfl_sync(0xffffffffu, my_num, i) & 0xFF;
                }
            }
            if (false)
#endif
            for (int t = 0; t < 34; t++) {
                if (R.root_key.tehai[t] == 0) continue;
                u8 th[34];
                for (int i = 0; i < 34; i++) th[i] = R.root_key.tehai[i];
                th[t] -= 1;
                int after = shanten_all_sig(s.T, sig_variant(base, t, -1, R.root_key.tehai[t]), len);
                SpCand& cd = cands[n_cands++];
                const int k5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
                cd.tile = (k5 >= 0 && ((R.root_key.akas >> k5) & 1) && R.root_key.tehai[t] == 1) ? T_5MR + k5 : t;
                cd.node = -1;
                cd.shanten_down = after - cur_shanten == 1;
                cd.t0 = cd.w0 = cd.e0 = 0.f;
                sp_required(s, c, len, th, R.root_key.wall, &cd.required, &cd.num_required);
            }
        } else {
            SpCand& cd = cands[n_cands++];
            cd.tile = T_UNK; cd.node = -1; cd.shanten_down = false; cd.t0 = cd.w0 = cd.e0 = 0.f;
            sp_required(s, c, len, R.root_key.tehai, R.root_key.wall, &cd.required, &cd.num_required);
        }
    } else {
        if (R.root == SP_NO_CHILD) return;
        if (can_discard) {  // calc.rs:203-253: one candidate per shanten-keeping discard of the root
            const u64 ri = s.G.einfo[sp_vid(s.G, R.root)];
            const int ne = sp_einfo_n(ri);
            const u32 eb = sp_einfo_begin(ri);
            for (int i = 0; i < ne && n_cands < 14; i++) {
                const u32 child = s.G.echild[eb + i];
                if (child == SP_NO_CHILD) continue;
                SpCand& cd = cands[n_cands++];
                cd.tile = s.G.emeta[eb + i] & 63;
                cd.node = (int)child;
                cd.shanten_down = false;
            }
        } else {
            SpCand& cd = cands[n_cands++];
            cd.tile = T_UNK; cd.node = (int)R.root; cd.shanten_down = false;
        }
#ifdef MJX_HOST_EMUL
        for (int i = 0; i < n_cands; i++) {
            SpCand& cd = cands[i];
            const int node = cd.node;
            cd.vid = sp_vid(s.G, (u32)node);
            const u64 ni = s.G.einfo[cd.vid];
            const int ne = sp_einfo_n(ni);
            const u32 eb = sp_einfo_begin(ni);
            u64 req = 0; int num = 0;
            for (int q = 0; q < ne; q++) {
                const u16 m = s.G.emeta[eb + q];
                req |= 1ull << deaka(m & 63);
                num += (m >> 6) & 7;
            }
            cd.required = req;
            cd.num_required = num & 0xFF;
            cd.t0 = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s.G, cd.vid, 0)[0]);
            cd.w0 = clamp01(sp_vals(s.G, cd.vid, 1)[0]);
            cd.e0 = fmaxf(sp_vals(s.G, cd.vid, 2)[0], 0.f);
        }
#else
        // the same, with the dependent loads of the candidates side by side: lane i fetches candidate i (dense id, edge
        // descriptor, first-turn values), the warp then walks each candidate's edges 32 at a time
        {
            u32 my_vid = 0, my_eb = 0;
            int my_ne = 0;
            float my_t0 = 0.f, my_w0 = 0.f, my_e0 = 0.f;
            if (s.lane < n_cands) {
                my_vid = sp_vid(s.G, (u32)cands[s.lane].node);
                const u64 ni = s.G.einfo[my_vid];
                my_ne = sp_einfo_n(ni); my_eb = sp_einfo_begin(ni);
                my_t0 = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s.G, my_vid, 0)[0]);
                my_w0 = clamp01(sp_vals(s.G, my_vid, 1)[0]);
                my_e0 = fmaxf(sp_vals(s.G, my_vid, 2)[0], 0.f);
            }
            for (int i = 0; i < n_cands; i++) {
                SpCand& cd = cands[i];
                cd.vid = __shfl_sync(0xffffffffu, my_vid, i);
                const int ne = __shfl_sync(0xffffffffu, my_ne, i);
                const u32 eb = __shfl_sync(0xffffffffu, my_eb, i);
                cd.t0 = __shfl_sync(0xffffffffu, my_t0, i);
                cd.w0 = __shfl_sync(0xffffffffu, my_w0, i);
                cd.e0 = __shfl_sync(0xffffffffu, my_e0, i);
                u32 lo = 0, hi = 0, num = 0;
                for (int q = s.lane; q < ne; q += 32) {
                    const u16 m = s.G.emeta[eb + q];
                    const int t = deaka(m & 63);
                    if (t < 32) lo |= 1u << t; else hi |= 1u << (t - 32);
                    num += (m >> 6) & 7;
                }
                lo = __reduce_or_sync(0xffffffffu, lo); hi = __reduce_or_sync(0xffffffffu, hi);
                num = __reduce_add_sync(0xffffffffu, num);
                cd.required = (u64)lo | ((u64)hi << 32);
                cd.num_required = (int)(num & 0xFF);
            }
        }
#endif
    }
    if (n_cands == 0) return;
    // `max_ev_table` is sorted descending (stable); index 0 = the maximum under the comparator
    int first = 0;
    for (int i = 1; i < n_cands; i++) if (sp_cand_cmp(cands[i], cands[first], has_values ? 0 : 3, has_values) > 0) first = i;
    if (R.after_riichi) cands[first].tile = R.last_self_tsumo;  // agent_helper.rs:588-590
    const float max_ev = has_values ? cands[first].e0 : 0.f;
    SPO_FILL(889, fminf(fmaxf(max_ev, 0.f), 100000.f) / 100000.f);
    SPO_FILL(890, fminf(fmaxf(max_ev, 0.f), 30000.f) / 30000.f);
    const bool cd_flag = R.cd_flag != 0;  // obs_repr.rs branches on cans.can_discard, not on the riichi-adjusted flag
    if (cd_flag) {
        for (int i = 0; i < n_cands; i++) {
            const int dt = deaka(cands[i].tile);
            const int r = 891 + (cands[i].shanten_down ? 34 : 0) + dt;
            MJX_FOR_TILES(s, t) { if ((cands[i].required >> t) & 1) obs_row[r * 34 + t] = 1.f; }
        }
        int best = 0;  // max_by(NotShantenDown) returns the LAST maximum
        for (int i = 1; i < n_cands; i++) if (sp_cand_cmp(cands[i], cands[best], 3, false) >= 0) best = i;
        SPO_ASSIGN(959, deaka(cands[best].tile), 1.f);
    } else {
        MJX_FOR_TILES(s, t) { if ((cands[first].required >> t) & 1) obs_row[960 * 34 + t] = 1.f; }
    }
    if (!has_values) return;
    if (!(cands[first].t0 > 0.f)) return;  // obs_repr.rs:645-653
    const float ev_scale = max_ev < 1.f ? 0.f : SP_FDIV(1.f, max_ev);
    const int n_emit = cd_flag ? n_cands : 1;
    for (int q = 0; q < n_emit; q++) {
        const SpCand& cd = cd_flag ? cands[q] : cands[first];
        const u32 vid = cd.vid;
#ifndef MJX_HOST_EMUL
        {   // lane = turn: the T <= 17 turns of a candidate are fetched side by side; take_while(p > 0) is a ballot
            const int turn = s.lane;
            float tp = 0.f, wp = 0.f, ev = 0.f;
            if (turn < T) {
                tp = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s.G, vid, 0)[turn]);
                wp = clamp01(sp_vals(s.G, vid, 1)[turn]);
                ev = fminf(SP_FMUL(fmaxf(sp_vals(s.G, vid, 2)[turn], 0.f), ev_scale), 1.f);
            }
            const unsigned stop = __ballot_sync(0xffffffffu, !(turn < T && tp > 0.f));  // never 0: T < 32
            const int n_turns = __ffs(stop) - 1;
            if (cd_flag) {
                if (turn < n_turns) {
                    const int tid = deaka(cd.tile);
                    obs_row[(961 + turn) * 34 + tid] = tp;
                    obs_row[(961 + SP_T_MAX + turn) * 34 + tid] = wp;
                    obs_row[(961 + 2 * SP_T_MAX + turn) * 34 + tid] = ev;
                }
            } else {
                for (int t = 0; t < n_turns; t++) {
                    const float a = __shfl_sync(0xffffffffu, tp, t), b = __shfl_sync(0xffffffffu, wp, t), c2 = __shfl_sync(0xffffffffu, ev, t);
                    SPO_FILL(961 + t, a);
                    SPO_FILL(961 + SP_T_MAX + t, b);
                    SPO_FILL(961 + 2 * SP_T_MAX + t, c2);
                }
            }
            continue;
        }
#endif
        for (int turn = 0; turn < T; turn++) {
            const float tp = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s.G, vid, 0)[turn]);
            if (!(tp > 0.f)) break;
            const float wp = clamp01(sp_vals(s.G, vid, 1)[turn]);
            const float ev = fminf(SP_FMUL(fmaxf(sp_vals(s.G, vid, 2)[turn], 0.f), ev_scale), 1.f);
            if (cd_flag) {
                const int tid = deaka(cd.tile);
                SPO_ASSIGN(961 + turn, tid, tp);
                SPO_ASSIGN(961 + SP_T_MAX + turn, tid, wp);
                SPO_ASSIGN(961 + 2 * SP_T_MAX + turn, tid, ev);
            } else {
                SPO_FILL(961 + turn, tp);
                SPO_FILL(961 + SP_T_MAX + turn, wp);
                SPO_FILL(961 + 2 * SP_T_MAX + turn, ev);
            }
        }
    }
#undef SPO_FILL
#undef SPO_ASSIGN
}

}  // namespace mjx
