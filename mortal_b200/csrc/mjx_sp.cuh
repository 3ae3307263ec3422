// mortal_b200 — single-player tables (obs v4 rows 889..1011) on device.
//
// Contract: libriichi state/agent_helper.rs:509-593 (single_player_tables) -> algo/sp/calc.rs
// (SPCalculator with calc_tegawari = calc_shanten_down = maximize_win_prob = false, which is how
// PlayerState calls it) -> state/obs_repr.rs:561-617, 632-692.
//
// The reference is a memoised recursion per observation (AHashMap<State, Rc<Values>> per shanten level).
// Here ALL observations of a step are solved together as one level-synchronous dynamic programme over the
// union of their (hand, wall) state DAGs:
//   slots      D3 W3 D2 W2 D1 W1 D0 W0   (D = 3n+2 hand choosing a shanten-keeping discard,
//                                          W = 3n+1 hand waiting for a shanten-lowering draw)
//   init       one warp per observation row: availability, parameters, root state -> its slot
//   expand     one launch per slot, one warp per state: lanes = the 34 tile ids (shanten of hand+-tile per
//              lane); children are interned in a global open-addressing hash table keyed by (row, state)
//   evaluate   slots in reverse, one warp per state: lane i owns turn i and accumulates over the state's
//              edges in the reference's iteration order with explicitly rounded f32 ops, so every
//              tenpai / win / EV value follows the same sequence of roundings as the Rust code
//   finalize   one warp per row: candidates, comparators, rows 889..1011 written into the obs tensor
// Work is balanced across the whole GPU at state granularity (a 15k-state hand costs as much as 30 small
// ones and is shared by all SMs), and there is no per-observation synchronisation.
#pragma once
#include "mjx_obs.cuh"

namespace mjx {

constexpr int SP_NODE_CHUNK = 16;   // state indices / work-list positions reserved per global atomic
constexpr int SP_EDGE_CHUNK = 128;  // edge slots reserved per global atomic
constexpr u32 SP_NO_OWNER = 0xFFFFFFFFu;  // edge_owner of a reserved-but-unused edge slot
constexpr int SP_T_MAX = 17;             // sp/mod.rs:42 MAX_TSUMOS_LEFT
constexpr int SP_SHANTEN_THRES = 3;      // calc.rs:13
constexpr int SP_MAX_TILES_LEFT = 34 * 4 - 1 - 13;  // calc.rs:14
constexpr int SP_SLOTS = 8;
constexpr u32 SP_NO_CHILD = 0xFFFFFFFFu;

#ifdef MJX_HOST_EMUL
#define SP_FMUL(a, b) ((a) * (b))
#define SP_FADD(a, b) ((a) + (b))
#define SP_FDIV(a, b) ((a) / (b))
#else
// never contracted into FMA: the reference rounds after every multiply and add
#define SP_FMUL(a, b) __fmul_rn((a), (b))
#define SP_FADD(a, b) __fadd_rn((a), (b))
#define SP_FDIV(a, b) __fdiv_rn((a), (b))
#endif

// sp/state.rs:10-21 (n_extra_tsumo is always 0 without tegawari)
struct alignas(8) SpKey {
    u8 tehai[34];
    u8 wall[34];
    u8 akas;      // bits 0-2 akas_in_hand, bits 3-5 akas_in_wall
    u8 pad_[3];
};
static_assert(sizeof(SpKey) == 72, "SpKey layout");

// 9 x 8-byte moves instead of 72 byte moves (the u8 arrays alone would only guarantee 1-byte alignment)
MJX_D void sp_key_copy(SpKey* dst, const SpKey* src) {
    const u64* a = reinterpret_cast<const u64*>(src);
    u64* b = reinterpret_cast<u64*>(dst);
#pragma unroll
    for (int i = 0; i < 9; i++) b[i] = a[i];
}

// What a state's expansion needs besides the key, carried from parent to child incrementally instead of being
// recomputed from the 34 counts: the shanten signatures of the hand (mjx_algo.cuh HandSig) and a Zobrist hash of the key.
struct alignas(8) SpSig {
    u32 idx[4];
    i8 kinds, pairs, kkinds, kpairs;
    u32 hash;
};
static_assert(sizeof(SpSig) == 24, "SpSig layout");
MJX_D HandSig sp_sig_hand(const SpSig& g) {
    HandSig h;
    for (int i = 0; i < 4; i++) h.idx[i] = g.idx[i];
    h.kinds = g.kinds; h.pairs = g.pairs; h.kkinds = g.kkinds; h.kpairs = g.kpairs;
    return h;
}
MJX_D SpSig sp_sig_make(const HandSig& h, u32 hash) {
    SpSig g;
    for (int i = 0; i < 4; i++) g.idx[i] = h.idx[i];
    g.kinds = (i8)h.kinds; g.pairs = (i8)h.pairs; g.kkinds = (i8)h.kkinds; g.kpairs = (i8)h.kpairs;
    g.hash = hash;
    return g;
}
// Zobrist terms: kind 0 = tehai[t] == c, 1 = wall[t] == c, 2 = the akas byte; the key hash is their XOR
MJX_D u32 sp_zob(int kind, int t, int c) {
    u32 x = ((u32)kind << 9) | ((u32)t << 3) | (u32)c;
    x = (x + 0x9E3779B9u) * 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
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
    i32 root;         // root node index, -1 if none
    float fallback_ev;  // obs_repr.rs:604-616
    SpKey root_key;
};

struct SpGlobal {
    SpRow* rows;       // [row_cap]
    SpKey* keys;       // [node_cap]
    SpSig* sigs;       // [node_cap]
    i32* node_row;     // [node_cap]
    float* vals;       // [node_cap][3][SP_T_MAX]
    u32* edge_begin;   // [node_cap]
    u8* n_edges;       // [node_cap]
    u32* edge_child;   // [edge_cap]
    u16* edge_meta;    // [edge_cap] tile (6 bits) | count << 6 | (no-yaku flag << 15, tenpai states only)
    u32* edge_owner;   // [edge_cap] node the edge belongs to
    float* leaf_scores;  // [score_cap][4] get_score of every winning draw of the tenpai (W0) states
    u32* hash;         // [hash_cap] node index + 1, 0 = empty
    i32* slot_list;    // [SP_SLOTS][slot_cap] node indices
    i32* slot_count;   // [SP_SLOTS]
    i32* counters;     // [0] nodes, [1] edges, [2] overflow flag, [3] overflow events (cumulative),
                       // [4],[5] edge range of the tenpai (W0) states
    i32 node_cap, edge_cap, hash_cap, slot_cap, score_cap;
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

// per-warp scratch (shared memory on device)
struct SpWarpScratch {
    float nts[SP_T_MAX];       // not_tsumo_prob row of the state (calc.rs:148-167)
    float tpn[SP_T_MAX];       // tsumo_prob[cnt-1][j] * not_tsumo[j] of the current edge
    float cv[3][SP_T_MAX];     // child values of the current edge
    float scores[40][4];       // get_score of each winning draw of a W0 state
    u8 score_ok[40];
    SpKey key;                   // the state being expanded, staged once per warp
    u8 ed_tile[40], ed_cnt[40];  // edge descriptors of the state being expanded
    u8 cand[40];                 // candidate tiles of the state being expanded, compacted
    u8 df[34];
    u8 pad_[2];
    i32 ed_n, ed_begin;
    // warp-private allocation chunks (device): indices are taken from the global counters SP_*_CHUNK at a time, so the
    // three hot counters see ~1/16 .. 1/128 of the atomics (same-address L2 atomics serialise, B300 guide 'Atomics')
    i32 a_node, a_node_end, a_pos, a_pos_end, a_edge, a_edge_end;
    i32 fill_from, fill_n, efill_from, efill_n, created, created_edges, bc_node, bc_pos;
};

struct SpCtx {
    SpGlobal G;
    Tables T;
    SpWarpScratch* ws;
    int lane;
};

#ifdef MJX_HOST_EMUL
#define SP_FOR_LANES(i, n) for (int i = 0; i < (n); i++)
#else
#define SP_FOR_LANES(i, n) for (int i = s.lane; i < (n); i += 32)
#endif

MJX_D u32 sp_key_hash_full(const SpKey& k) {
    u32 h = sp_zob(2, 0, k.akas);
    for (int t = 0; t < 34; t++) h ^= sp_zob(0, t, k.tehai[t]) ^ sp_zob(1, t, k.wall[t]);
    return h;
}
// table slot hash of (row, key): the key's Zobrist hash mixed with the row
MJX_D u32 sp_hash_key(int row, u32 key_hash) {
    u32 x = key_hash ^ ((u32)row * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

// `a` lives in the global arena and may have been written by another SM during this launch: read it through
// L2 (ld.global.cg) — an L1 line shared with a neighbouring, older node could otherwise serve stale bytes.
MJX_D bool sp_key_eq(const SpKey& a, const SpKey& b) {
    const u32* x = reinterpret_cast<const u32*>(&a);
    const u32* y = reinterpret_cast<const u32*>(&b);
    bool eq = true;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SpKey) / 4); i++) {
#ifdef MJX_HOST_EMUL
        eq &= x[i] == y[i];
#else
        eq &= __ldcg(x + i) == y[i];
#endif
    }
    return eq;
}
MJX_D int sp_ld_row(const i32* p) {
#ifdef MJX_HOST_EMUL
    return *p;
#else
    return __ldcg(p);
#endif
}

MJX_D void sp_set_overflow(SpCtx& s) { s.G.counters[2] = 1; }

// allocate a node in `slot`; executed by one lane. Returns -1 on overflow.
MJX_DN int sp_new_node(SpCtx& s, int row, const SpKey& key, const SpSig& sig, int slot) {
#ifdef MJX_HOST_EMUL
    int idx = s.G.counters[0]++;
    int pos = s.G.slot_count[slot]++;
#else
    int idx = atomicAdd(&s.G.counters[0], 1);
    int pos = atomicAdd(&s.G.slot_count[slot], 1);
    atomicAdd(&s.G.counters[6], 1);  // states actually created (counters[0] also counts abandoned chunk tails)
#endif
    if (idx >= s.G.node_cap || pos >= s.G.slot_cap) { sp_set_overflow(s); return -1; }
    sp_key_copy(&s.G.keys[idx], &key);
    {
        const u64* a = reinterpret_cast<const u64*>(&sig);
        u64* b = reinterpret_cast<u64*>(&s.G.sigs[idx]);
        b[0] = a[0]; b[1] = a[1]; b[2] = a[2];
    }
    s.G.node_row[idx] = row;
    s.G.n_edges[idx] = 0;
    s.G.edge_begin[idx] = 0;
    s.G.slot_list[(size_t)slot * s.G.slot_cap + pos] = idx;
    return idx;
}

// find-or-insert (row, key) in `slot`. May be called by several lanes of a warp at once (different keys).
// Lock-free without spinning: the node is allocated and written first, then published with one CAS; if another
// thread published the same state in the meantime its node wins and ours is simply never referenced.
MJX_DN int sp_intern(SpCtx& s, int row, const SpKey& key, const SpSig& sig, int slot) {
    const u32 mask = (u32)s.G.hash_cap - 1;
    const u32 hv = sp_hash_key(row, sig.hash);
    const u32 tag = (hv >> 24) << 24;  // 8-bit tag kept beside the 24-bit index: mismatches never touch the key array
    u32 h = hv & mask;
    int mine = -1;
    for (int probe = 0; probe < s.G.hash_cap; probe++, h = (h + 1) & mask) {
#ifdef MJX_HOST_EMUL
        u32 cur = s.G.hash[h];
        if (cur == 0) {
            int idx = sp_new_node(s, row, key, sig, slot);
            if (idx < 0) return -1;
            s.G.hash[h] = ((u32)idx + 1) | tag;
            return idx;
        }
#else
        u32 cur = __ldcg(&s.G.hash[h]);
        if (cur == 0) {
            if (mine < 0) {
                mine = sp_new_node(s, row, key, sig, slot);
                if (mine < 0) return -1;
                __threadfence();
            }
            const u32 prev = atomicCAS(&s.G.hash[h], 0u, ((u32)mine + 1) | tag);
            if (prev == 0) return mine;
            cur = prev;
        }
#endif
        if ((cur & 0xFF000000u) != tag) continue;
        const int ci = (int)(cur & 0x00FFFFFFu) - 1;
        if (sp_ld_row(s.G.node_row + ci) == row && sp_key_eq(s.G.keys[ci], key)) return ci;
    }
    sp_set_overflow(s);
    return -1;
}

#ifndef MJX_HOST_EMUL
// Device interning in two phases so that new states are allocated per warp, not per lane.
// Phase 1: read-only probe. Returns the state's index, or -1 with `h` left at the first empty hash slot seen.
MJX_DN int sp_lookup(const SpCtx& s, int row, const SpKey& key, u32 hv, u32& h) {
    const u32 mask = (u32)s.G.hash_cap - 1;
    const u32 tag = (hv >> 24) << 24;
    h = hv & mask;
    for (int probe = 0; probe < s.G.hash_cap; probe++, h = (h + 1) & mask) {
        const u32 cur = __ldcg(&s.G.hash[h]);
        if (cur == 0) return -1;
        if ((cur & 0xFF000000u) != tag) continue;
        const int ci = (int)(cur & 0x00FFFFFFu) - 1;
        if (sp_ld_row(s.G.node_row + ci) == row && sp_key_eq(s.G.keys[ci], key)) return ci;
    }
    return -1;
}
// Phase 2: publish the already written state `mine` starting at hash slot `h`. If another warp published the same
// state in the meantime its index is returned and `lost` is set (ours becomes a hole in the work list).
MJX_DN int sp_publish(SpCtx& s, int row, const SpKey& key, u32 hv, u32 h, int mine, bool& lost) {
    const u32 mask = (u32)s.G.hash_cap - 1;
    const u32 tag = (hv >> 24) << 24;
    lost = false;
    for (int probe = 0; probe < s.G.hash_cap; probe++, h = (h + 1) & mask) {
        u32 cur = __ldcg(&s.G.hash[h]);
        if (cur == 0) {
            const u32 prev = atomicCAS(&s.G.hash[h], 0u, ((u32)mine + 1) | tag);
            if (prev == 0) return mine;
            cur = prev;
        }
        if ((cur & 0xFF000000u) != tag) continue;
        const int ci = (int)(cur & 0x00FFFFFFu) - 1;
        if (sp_ld_row(s.G.node_row + ci) == row && sp_key_eq(s.G.keys[ci], key)) { lost = true; return ci; }
    }
    sp_set_overflow(s);
    lost = true;
    return -1;
}
// Warp-collective: reserve `count` consecutive state indices and work-list positions of `slot` from the warp's chunks,
// refilling a chunk from its global counter when it runs short (the abandoned remainder stays a run of holes: the
// positions were pre-filled with -1 when the chunk was taken). Returns false on arena overflow.
MJX_DN bool sp_reserve(SpCtx& s, int slot, int count, int& node_base, int& pos_base) {
    SpWarpScratch& ws = *s.ws;
    __syncwarp();
    if (s.lane == 0) {
        if (ws.a_node_end - ws.a_node < count) {
            const int n = max(SP_NODE_CHUNK, count);
            ws.a_node = atomicAdd(&s.G.counters[0], n);
            ws.a_node_end = ws.a_node + n;
        }
        ws.fill_n = 0;
        if (ws.a_pos_end - ws.a_pos < count) {
            const int n = max(SP_NODE_CHUNK, count);
            ws.a_pos = atomicAdd(&s.G.slot_count[slot], n);
            ws.a_pos_end = ws.a_pos + n;
            ws.fill_from = ws.a_pos; ws.fill_n = n;
        }
        ws.bc_node = ws.a_node; ws.bc_pos = ws.a_pos;
        ws.a_node += count; ws.a_pos += count;
    }
    __syncwarp();
    node_base = ws.bc_node; pos_base = ws.bc_pos;
    const int ff = ws.fill_from, fn = ws.fill_n;
    i32* list = s.G.slot_list + (size_t)slot * s.G.slot_cap;
    for (int i = s.lane; i < fn; i += 32) if (ff + i < s.G.slot_cap) list[ff + i] = -1;
    __syncwarp();
    if (node_base + count > s.G.node_cap || pos_base + count > s.G.slot_cap) { sp_set_overflow(s); return false; }
    return true;
}
#endif

MJX_HD bool sp_slot_is_w(int slot) { return (slot & 1) != 0; }
MJX_HD int sp_slot_shanten(int slot) { return 3 - (slot >> 1); }
MJX_D float* sp_vals(const SpCtx& s, int node, int which) { return s.G.vals + ((size_t)node * 3 + which) * SP_T_MAX; }

// Expand one state (one warp): edges = shanten-lowering draws (W) or shanten-keeping discards (D).
MJX_DN void sp_expand(SpCtx& s, const Ctx& c, int node, int slot) {
    const bool is_w = sp_slot_is_w(slot);
    const int k = sp_slot_shanten(slot);
    const bool leaf = is_w && k == 0;
    const int row = s.G.node_row[node];
    SpWarpScratch& ws = *s.ws;
    MJX_SYNCWARP();
#ifdef MJX_HOST_EMUL
    sp_key_copy(&ws.key, &s.G.keys[node]);
#else
    if (s.lane < 9) reinterpret_cast<u64*>(&ws.key)[s.lane] = reinterpret_cast<const u64*>(&s.G.keys[node])[s.lane];
#endif
    const SpSig sg = s.G.sigs[node];
    MJX_SYNCWARP();
    const SpKey& key = ws.key;
    const int len = s.G.rows[row].tehai_len_div3;
    const HandSig base = sp_sig_hand(sg);
    // candidate tiles: still in the wall (W) / held (D); only those need a shanten evaluation
    const u8* cnts = is_w ? key.wall : key.tehai;
    const u64 cand = tile_mask(c, [&](int t) { return cnts[t] != 0; });
    auto effective = [&](int t) {
        if (is_w) return shanten_all_sig(s.T, sig_variant(base, t, +1, key.tehai[t]), len) - k == -1;
        return shanten_all_sig(s.T, sig_variant(base, t, -1, key.tehai[t]), len) == k;
    };
    u64 eff = 0;
#ifdef MJX_HOST_EMUL
    for (int t = 0; t < 34; t++) if (((cand >> t) & 1) && effective(t)) eff |= 1ull << t;
#else
    {   // lane i evaluates the i-th candidate: one pass unless more than 32 kinds qualify
        for (int t = s.lane; t < 34; t += 32)
            if ((cand >> t) & 1) ws.cand[mjx_popcll(cand & ((1ull << t) - 1))] = (u8)t;
        __syncwarp();
        const int n_c = mjx_popcll(cand);
        for (int i0 = 0; i0 < n_c; i0 += 32) {
            const int i = i0 + s.lane;
            const int t = i < n_c ? (int)ws.cand[i] : -1;
            const bool ok = t >= 0 && effective(t);
            const unsigned lo = (ok && t < 32) ? 1u << t : 0u, hi = (ok && t >= 32) ? 1u << (t - 32) : 0u;
            eff |= (u64)__reduce_or_sync(0xffffffffu, lo) | ((u64)__reduce_or_sync(0xffffffffu, hi) << 32);
        }
    }
#endif
    // edge descriptors in tile order; an effective 5 whose aka is still in the wall splits in two
    // (sp/state.rs:160-176); a discarded 5 is the aka only when it is the last 5 in hand (state.rs:127-132).
    // Every effective tile writes its own descriptors at the offset its rank gives.
    u64 split = 0;
    if (is_w)
        for (int s5 = 0; s5 < 3; s5++) {
            const int t5 = 4 + 9 * s5;
            if (((eff >> t5) & 1) && ((key.akas >> (3 + s5)) & 1) && key.wall[t5] >= 2) split |= 1ull << t5;
        }
    const int ne_all = mjx_popcll(eff) + mjx_popcll(split);
    MJX_FOR_TILES(c, t) {
        if (!((eff >> t) & 1)) continue;
        const u64 below = (1ull << t) - 1;
        const int off = mjx_popcll(eff & below) + mjx_popcll(split & below);
        const int suit5 = (t == T_5M || t == T_5P || t == T_5S) ? t / 9 : -1;
        if (is_w) {
            const int count = key.wall[t];
            if (suit5 >= 0 && ((key.akas >> (3 + suit5)) & 1)) {
                int o = off;
                if (count >= 2) { ws.ed_tile[o] = (u8)t; ws.ed_cnt[o] = (u8)(count - 1); o++; }
                ws.ed_tile[o] = (u8)(T_5MR + suit5); ws.ed_cnt[o] = 1;
            } else {
                ws.ed_tile[off] = (u8)t; ws.ed_cnt[off] = (u8)count;
            }
        } else {
            int tile = t;
            if (suit5 >= 0 && ((key.akas >> suit5) & 1) && key.tehai[t] == 1) tile = T_5MR + suit5;
            ws.ed_tile[off] = (u8)tile; ws.ed_cnt[off] = 0;
        }
    }
    if (MJX_IS_L0(c)) {
        int ne = ne_all;
#ifdef MJX_HOST_EMUL
        int eb = s.G.counters[1]; s.G.counters[1] += ne;
#else
        ws.efill_n = 0;
        if (ws.a_edge_end - ws.a_edge < ne) {  // take a new chunk of edge slots; the remainder of the old one stays unused
            // the tenpai level's edge range is walked edge by edge by k_sp_score: keep its unused tails short
            const int n = max(leaf ? 32 : SP_EDGE_CHUNK, ne);
            ws.a_edge = atomicAdd(&s.G.counters[1], n);
            ws.a_edge_end = ws.a_edge + n;
            ws.efill_from = ws.a_edge; ws.efill_n = n;
        }
        int eb = ws.a_edge;
        ws.a_edge += ne;
        ws.created_edges += ne;
#endif
        if (eb + ne > s.G.edge_cap) { sp_set_overflow(s); ne = 0; eb = 0; }
        ws.ed_n = ne; ws.ed_begin = eb;
        s.G.edge_begin[node] = (u32)eb;
        s.G.n_edges[node] = (u8)ne;
    }
    MJX_SYNCWARP();
#ifndef MJX_HOST_EMUL
    {   // unused edge slots of a fresh chunk must read as "no edge" for k_sp_score, which walks the edge range
        const int ff = ws.efill_from, fn = ws.efill_n;
        for (int i = s.lane; i < fn; i += 32) if (ff + i < s.G.edge_cap) s.G.edge_owner[ff + i] = SP_NO_OWNER;
        __syncwarp();
    }
#endif
    // one edge per lane: build the child state (key, signatures, hash: all incremental) and intern it
    const int ne = ws.ed_n, eb = ws.ed_begin;
    auto child_of = [&](int e, SpKey& ck, SpSig& cs) {
        const int tile = ws.ed_tile[e], t = deaka(tile);
        const int suit5 = is_aka(tile) ? tile - T_5MR : -1;
        sp_key_copy(&ck, &key);
        const int c0 = key.tehai[t];
        u32 h = sg.hash ^ sp_zob(0, t, c0);
        if (is_w) {
            ck.tehai[t] += 1;
            ck.wall[t] -= 1;
            if (suit5 >= 0) ck.akas = (u8)((ck.akas | (1 << suit5)) & ~(1 << (3 + suit5)));
            h ^= sp_zob(0, t, c0 + 1) ^ sp_zob(1, t, key.wall[t]) ^ sp_zob(1, t, key.wall[t] - 1);
        } else {
            ck.tehai[t] -= 1;
            if (suit5 >= 0) ck.akas = (u8)(ck.akas & ~(1 << suit5));
            h ^= sp_zob(0, t, c0 - 1);
        }
        if (ck.akas != key.akas) h ^= sp_zob(2, 0, key.akas) ^ sp_zob(2, 0, ck.akas);
        cs = sp_sig_make(sig_variant(base, t, is_w ? +1 : -1, c0), h);
    };
#ifdef MJX_HOST_EMUL
    for (int e = 0; e < ne; e++) {
        u32 child = SP_NO_CHILD;
        if (!leaf) {
            SpKey ck; SpSig cs;
            child_of(e, ck, cs);
            int ci = sp_intern(s, row, ck, cs, slot + 1);
            child = ci < 0 ? SP_NO_CHILD : (u32)ci;
        }
        s.G.edge_child[eb + e] = child;
        s.G.edge_meta[eb + e] = (u16)(ws.ed_tile[e] | (ws.ed_cnt[e] << 6));
        s.G.edge_owner[eb + e] = (u32)node;
    }
#else
    for (int e0 = 0; e0 < ne; e0 += 32) {
        const int e = e0 + s.lane;
        const bool active = e < ne;
        const int my_tile = active ? ws.ed_tile[e] : 0, my_cnt = active ? ws.ed_cnt[e] : 0;
        u32 child = SP_NO_CHILD;
        if (!leaf) {
            SpKey ck; SpSig cs;
            u32 hv = 0, h = 0;
            int found = -1;
            if (active) {
                child_of(e, ck, cs);
                hv = sp_hash_key(row, cs.hash);
                found = sp_lookup(s, row, ck, hv, h);
            }
            const bool need = active && found < 0;
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int node_base = 0, pos_base = 0;
                const bool ok = sp_reserve(s, slot + 1, __popc(m), node_base, pos_base);
                if (need && ok) {
                    const int rank = __popc(m & ((1u << s.lane) - 1));
                    const int idx = node_base + rank, pos = pos_base + rank;
                    sp_key_copy(&s.G.keys[idx], &ck);
                    {
                        const u64* a = reinterpret_cast<const u64*>(&cs);
                        u64* b = reinterpret_cast<u64*>(&s.G.sigs[idx]);
                        b[0] = a[0]; b[1] = a[1]; b[2] = a[2];
                    }
                    s.G.node_row[idx] = row;
                    s.G.n_edges[idx] = 0;
                    s.G.edge_begin[idx] = 0;
                    i32* list = s.G.slot_list + (size_t)(slot + 1) * s.G.slot_cap;
                    list[pos] = idx;
                    __threadfence();
                    bool lost;
                    found = sp_publish(s, row, ck, hv, h, idx, lost);
                    if (lost) list[pos] = -1;
                }
                if (s.lane == 0) ws.created += __popc(m);
            }
            child = found < 0 ? SP_NO_CHILD : (u32)found;
        }
        if (active) {
            s.G.edge_child[eb + e] = child;
            s.G.edge_meta[eb + e] = (u16)(my_tile | (my_cnt << 6));
            s.G.edge_owner[eb + e] = (u32)node;
        }
    }
#endif
    MJX_SYNCWARP();
}

// calc.rs:640-758 for one winning draw; executed by one lane. Returns false when there is no yaku.
MJX_DN bool sp_get_score(const SpCtx& s, const SpRow& P, const SpKey& key, int win_tile, float* scores) {
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
    q.chis = P.melds; q.pons = P.melds + 4; q.minkans = P.melds + 8; q.ankans = P.melds + 12;
    q.n_chis = P.n_chis; q.n_pons = P.n_pons; q.n_minkans = P.n_minkans; q.n_ankans = P.n_ankans;
    q.bakaze = P.bakaze; q.jikaze = P.jikaze; q.winning_tile = wid; q.is_ron = false; q.is_menzen = P.is_menzen != 0;
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

// calc.rs:447-561 draw_without_tegawari_slow for one W-state at shanten k (one warp, lane i = turn i)
// stage: score every winning draw of every tenpai state, ONE THREAD PER DRAW (calc.rs:478-479 get_score).
// Full lane utilisation for the branchy agari evaluation; the per-turn accumulation happens in sp_eval_w<true>.
MJX_DN void sp_score_edge(const SpCtx& s, int e) {
    if (s.G.edge_owner[e] == SP_NO_OWNER) return;  // reserved but unused edge slot
    const int node = (int)s.G.edge_owner[e];
    const SpRow& P = s.G.rows[s.G.node_row[node]];
    float sc[4];
    const u16 m = s.G.edge_meta[e];
    SpKey key;
    sp_key_copy(&key, &s.G.keys[node]);
    const bool ok = sp_get_score(s, P, key, m & 63, sc);
    const int le = e - s.G.counters[4];
    if (le >= s.G.score_cap) s.G.counters[2] = 1;  // score arena too small for this step: reported as an overflow
    if (!ok) s.G.edge_meta[e] = (u16)(m | 0x8000);
    else if (le >= 0 && le < s.G.score_cap) {
        float* o = s.G.leaf_scores + (size_t)le * 4;
        o[0] = sc[0]; o[1] = sc[1]; o[2] = sc[2]; o[3] = sc[3];
    }
}

template <bool LEAF>
MJX_DN void sp_eval_w(SpCtx& s, const Ctx& c, int node, int k_rt) {
    const int k = LEAF ? 0 : k_rt;
    const int row = s.G.node_row[node];
    const SpRow& P = s.G.rows[row];
    const int T = P.T, n_left = P.n_left;
    const int ne = s.G.n_edges[node];
    const u32 eb = s.G.edge_begin[node];
    SpWarpScratch& ws = *s.ws;
    int sum_required = 0;
    for (int e = 0; e < ne; e++) sum_required += (s.G.edge_meta[eb + e] >> 6) & 7;
    sum_required &= 0xFF;
    // not_tsumo_prob_table[sum_required][j], recomputed with the table's own recurrence (calc.rs:158-165)
    SP_FOR_LANES(j, T) {
        float v = 0.f;
        const int i0 = sum_required;
        if (i0 <= n_left && i0 <= SP_MAX_TILES_LEFT) {
            const int lim = min(T - 1, n_left - i0);
            if (j <= lim) {
                v = 1.f;
                for (int q = 0; q < j; q++) v = SP_FDIV(SP_FMUL(v, (float)(n_left - i0 - q)), (float)(n_left - q));
            }
        }
        ws.nts[j] = v;
    }
    if (k == 0) {
        // scores were produced by sp_score_edge; stage them for the turn lanes
        const int le0 = (int)eb - s.G.counters[4];
        SP_FOR_LANES(e, ne) {
            const bool ok = !(s.G.edge_meta[eb + e] & 0x8000) && le0 + e >= 0 && le0 + e < s.G.score_cap;
            ws.score_ok[e] = ok ? 1 : 0;
            if (ok) for (int q = 0; q < 4; q++) ws.scores[e][q] = s.G.leaf_scores[(size_t)(le0 + e) * 4 + q];
        }
    }
    MJX_SYNCWARP();
    float tenpai = 0.f, win = 0.f, ev = 0.f;  // lane-private accumulators (emulation: see below)
#ifdef MJX_HOST_EMUL
    float a_t[SP_T_MAX] = {0}, a_w[SP_T_MAX] = {0}, a_v[SP_T_MAX] = {0};
#endif
    const bool assume_riichi = P.is_menzen && P.prefer_riichi;
    for (int e = 0; e < ne; e++) {
        const int cnt = (s.G.edge_meta[eb + e] >> 6) & 7;
        if (k == 0 && !ws.score_ok[e]) continue;  // uniform
        const u32 child = s.G.edge_child[eb + e];
        if (k > 0 && child == SP_NO_CHILD) continue;  // only after an overflow
        MJX_SYNCWARP();
        SP_FOR_LANES(j, T) {
            ws.tpn[j] = SP_FMUL(SP_FDIV((float)cnt, (float)(n_left - j)), ws.nts[j]);
            if (k > 0) {
                ws.cv[0][j] = sp_vals(s, (int)child, 0)[j];
                ws.cv[1][j] = sp_vals(s, (int)child, 1)[j];
                ws.cv[2][j] = sp_vals(s, (int)child, 2)[j];
            }
        }
        MJX_SYNCWARP();
        SP_FOR_LANES(i, T) {
#ifdef MJX_HOST_EMUL
            tenpai = a_t[i]; win = a_w[i]; ev = a_v[i];
#endif
            const float m = ws.nts[i];
            if (m != 0.f) {
                for (int j = i; j < T; j++) {
                    if (ws.nts[j] == 0.f) break;
                    const float prob = SP_FDIV(ws.tpn[j], m);
                    if (k == 0) {
                        const int han_plus = (assume_riichi && P.calc_double_riichi && i == 0) + (assume_riichi && j == i) +
                                             (P.calc_haitei && j == T - 1);
                        win = SP_FADD(win, prob);
                        ev = SP_FADD(ev, SP_FMUL(prob, ws.scores[e][han_plus]));
                    } else {
                        if (k == 1) tenpai = SP_FADD(tenpai, prob);
                        if (j < T - 1) {
                            if (k > 1) tenpai = SP_FADD(tenpai, SP_FMUL(prob, ws.cv[0][j + 1]));
                            win = SP_FADD(win, SP_FMUL(prob, ws.cv[1][j + 1]));
                            ev = SP_FADD(ev, SP_FMUL(prob, ws.cv[2][j + 1]));
                        }
                    }
                }
            }
#ifdef MJX_HOST_EMUL
            a_t[i] = tenpai; a_w[i] = win; a_v[i] = ev;
#endif
        }
    }
    MJX_SYNCWARP();
    SP_FOR_LANES(i, T) {
#ifdef MJX_HOST_EMUL
        tenpai = a_t[i]; win = a_w[i]; ev = a_v[i];
#endif
        sp_vals(s, node, 0)[i] = tenpai;
        sp_vals(s, node, 1)[i] = win;
        sp_vals(s, node, 2)[i] = ev;
    }
    MJX_SYNCWARP();
}

#ifndef MJX_HOST_EMUL
// Device form of sp_eval_w: TWO states per warp, one per half-warp (turn i = sub-lane; a state has at most 17
// turns, sub-lane 15 also carries turn 16). Every lane performs exactly the float operations sp_eval_w performs for
// its turn, in the same order, so the values are bit-identical; the halves only share the instruction stream.
struct SpEvalScratch {
    float nts[2][SP_T_MAX], tpn[2][SP_T_MAX], cv[2][3][SP_T_MAX];
    float scores[2][40][4];
    u8 score_ok[2][40];
};

template <bool LEAF>
MJX_DN void sp_eval_w2(SpCtx& s, SpEvalScratch& es, int node, int k_rt) {
    const int k = LEAF ? 0 : k_rt;
    const int hw = s.lane >> 4, l = s.lane & 15;
    const bool live = node >= 0;
    const SpRow* P = live ? &s.G.rows[s.G.node_row[node]] : nullptr;
    const int T = live ? P->T : 0, n_left = live ? P->n_left : 0;
    const int ne = live ? s.G.n_edges[node] : 0;
    const u32 eb = live ? s.G.edge_begin[node] : 0;
    float* nts = es.nts[hw];
    float* tpn = es.tpn[hw];
    int sum_required = 0;
    for (int e = 0; e < ne; e++) sum_required += (s.G.edge_meta[eb + e] >> 6) & 7;
    sum_required &= 0xFF;
    for (int j = l; j < T; j += 16) {  // not_tsumo_prob_table[sum_required][j] (calc.rs:158-165)
        float v = 0.f;
        const int i0 = sum_required;
        if (i0 <= n_left && i0 <= SP_MAX_TILES_LEFT) {
            const int lim = min(T - 1, n_left - i0);
            if (j <= lim) {
                v = 1.f;
                for (int q = 0; q < j; q++) v = SP_FDIV(SP_FMUL(v, (float)(n_left - i0 - q)), (float)(n_left - q));
            }
        }
        nts[j] = v;
    }
    if (k == 0) {
        const int le0 = (int)eb - s.G.counters[4];
        for (int e = l; e < ne; e += 16) {
            const bool ok = !(s.G.edge_meta[eb + e] & 0x8000) && le0 + e >= 0 && le0 + e < s.G.score_cap;
            es.score_ok[hw][e] = ok ? 1 : 0;
            if (ok) for (int q = 0; q < 4; q++) es.scores[hw][e][q] = s.G.leaf_scores[(size_t)(le0 + e) * 4 + q];
        }
    }
    __syncwarp();
    float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f};  // (tenpai, win, ev) of turn l and of turn 16
    const bool assume_riichi = live && P->is_menzen && P->prefer_riichi;
    const bool dbl = live && P->calc_double_riichi, haitei = live && P->calc_haitei;
    const int ne_max = max(ne, __shfl_xor_sync(0xffffffffu, ne, 16));
    for (int e = 0; e < ne_max; e++) {
        bool skip = e >= ne;
        int cnt = 0;
        u32 child = SP_NO_CHILD;
        if (!skip) {
            cnt = (s.G.edge_meta[eb + e] >> 6) & 7;
            if (k == 0 && !es.score_ok[hw][e]) skip = true;
            child = s.G.edge_child[eb + e];
            if (k > 0 && child == SP_NO_CHILD) skip = true;  // only after an overflow
        }
        __syncwarp();
        if (!skip)
            for (int j = l; j < T; j += 16) {
                tpn[j] = SP_FMUL(SP_FDIV((float)cnt, (float)(n_left - j)), nts[j]);
                if (k > 0) {
                    es.cv[hw][0][j] = sp_vals(s, (int)child, 0)[j];
                    es.cv[hw][1][j] = sp_vals(s, (int)child, 1)[j];
                    es.cv[hw][2][j] = sp_vals(s, (int)child, 2)[j];
                }
            }
        __syncwarp();
        if (!skip) {
            auto turn = [&](int i, float* a) {
                float tenpai = a[0], win = a[1], ev = a[2];
                const float m = nts[i];
                if (m != 0.f) {
                    for (int j = i; j < T; j++) {
                        if (nts[j] == 0.f) break;
                        const float prob = SP_FDIV(tpn[j], m);
                        if (k == 0) {
                            const int han_plus = (assume_riichi && dbl && i == 0) + (assume_riichi && j == i) + (haitei && j == T - 1);
                            win = SP_FADD(win, prob);
                            ev = SP_FADD(ev, SP_FMUL(prob, es.scores[hw][e][han_plus]));
                        } else {
                            if (k == 1) tenpai = SP_FADD(tenpai, prob);
                            if (j < T - 1) {
                                if (k > 1) tenpai = SP_FADD(tenpai, SP_FMUL(prob, es.cv[hw][0][j + 1]));
                                win = SP_FADD(win, SP_FMUL(prob, es.cv[hw][1][j + 1]));
                                ev = SP_FADD(ev, SP_FMUL(prob, es.cv[hw][2][j + 1]));
                            }
                        }
                    }
                }
                a[0] = tenpai; a[1] = win; a[2] = ev;
            };
            if (l < T) turn(l, a0);
            if (T > 16 && l == 15) turn(16, a1);
        }
    }
    __syncwarp();
    if (live) {
        if (l < T) for (int q = 0; q < 3; q++) sp_vals(s, node, q)[l] = a0[q];
        if (T > 16 && l == 15) for (int q = 0; q < 3; q++) sp_vals(s, node, q)[16] = a1[q];
    }
    __syncwarp();
}
#endif

// calc.rs:563-637 discard_slow for one D-state (one warp, lane i = turn i)
MJX_DN void sp_eval_d(SpCtx& s, const Ctx& c, int node) {
    const int T = s.G.rows[s.G.node_row[node]].T;
    const int ne = s.G.n_edges[node];
    const u32 eb = s.G.edge_begin[node];
    SP_FOR_LANES(i, T) {
        const float FMIN = -3.40282347e+38f;
        float bt = FMIN, bw = FMIN, bv = FMIN;
        int best_tile = T_UNK;
        i32 best_value = (i32)0x80000000;
        for (int e = 0; e < ne; e++) {
            const u32 child = s.G.edge_child[eb + e];
            if (child == SP_NO_CHILD) continue;
            const int tile = s.G.edge_meta[eb + e] & 63;
            const float v = sp_vals(s, (int)child, 2)[i];
            const i32 value = (i32)v;  // finite and < 2^31 here; Rust `as i32` truncates the same way
            if (value > best_value || (value == best_value && cmp_discard_priority(tile, best_tile) > 0)) {
                bt = sp_vals(s, (int)child, 0)[i]; bw = sp_vals(s, (int)child, 1)[i]; bv = v;
                best_value = value; best_tile = tile;
            }
        }
        sp_vals(s, node, 0)[i] = bt;
        sp_vals(s, node, 1)[i] = bw;
        sp_vals(s, node, 2)[i] = bv;
    }
    MJX_SYNCWARP();
}

#ifndef MJX_HOST_EMUL
// Device form of sp_eval_d: two D-states per warp, one per half-warp (turn i = sub-lane, sub-lane 15 also carries turn 16).
MJX_DN void sp_eval_d2(SpCtx& s, int node) {
    if (node < 0) return;
    const int l = s.lane & 15;
    const int T = s.G.rows[s.G.node_row[node]].T;
    const int ne = s.G.n_edges[node];
    const u32 eb = s.G.edge_begin[node];
    for (int pass = 0; pass < 2; pass++) {
        const int i = pass == 0 ? l : 16;
        if (pass == 0 ? i >= T : !(T > 16 && l == 15)) continue;
        const float FMIN = -3.40282347e+38f;
        float bt = FMIN, bw = FMIN, bv = FMIN;
        int best_tile = T_UNK;
        i32 best_value = (i32)0x80000000;
        for (int e = 0; e < ne; e++) {
            const u32 child = s.G.edge_child[eb + e];
            if (child == SP_NO_CHILD) continue;
            const int tile = s.G.edge_meta[eb + e] & 63;
            const float v = sp_vals(s, (int)child, 2)[i];
            const i32 value = (i32)v;  // finite and < 2^31 here; Rust `as i32` truncates the same way
            if (value > best_value || (value == best_value && cmp_discard_priority(tile, best_tile) > 0)) {
                bt = sp_vals(s, (int)child, 0)[i]; bw = sp_vals(s, (int)child, 1)[i]; bv = v;
                best_value = value; best_tile = tile;
            }
        }
        sp_vals(s, node, 0)[i] = bt;
        sp_vals(s, node, 1)[i] = bw;
        sp_vals(s, node, 2)[i] = bv;
    }
}
#endif

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
        s.ws->df[t] = (u8)f;
    }
    MJX_END_TILES(s);
    const u8* df = s.ws->df;
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
        R.root = -1;
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
            sp_key_copy(&R.root_key, &root);
            if (R.has_values) {
                const int slot = 2 * (3 - cur_shanten) + (R.can_discard ? 0 : 1);
                R.root = sp_new_node(s, row, root, sp_sig_make(hand_sig(root.tehai), sp_key_hash_full(root)), slot);
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
        if (R.root < 0) return;
        if (can_discard) {  // calc.rs:203-253: one candidate per shanten-keeping discard of the root
            const int ne = s.G.n_edges[R.root];
            const u32 eb = s.G.edge_begin[R.root];
            for (int i = 0; i < ne && n_cands < 14; i++) {
                const u32 child = s.G.edge_child[eb + i];
                if (child == SP_NO_CHILD) continue;
                SpCand& cd = cands[n_cands++];
                cd.tile = s.G.edge_meta[eb + i] & 63;
                cd.node = (int)child;
                cd.shanten_down = false;
            }
        } else {
            SpCand& cd = cands[n_cands++];
            cd.tile = T_UNK; cd.node = R.root; cd.shanten_down = false;
        }
        for (int i = 0; i < n_cands; i++) {
            SpCand& cd = cands[i];
            const int node = cd.node;
            const int ne = s.G.n_edges[node];
            const u32 eb = s.G.edge_begin[node];
            u64 req = 0; int num = 0;
            for (int q = 0; q < ne; q++) {
                const u16 m = s.G.edge_meta[eb + q];
                req |= 1ull << deaka(m & 63);
                num += (m >> 6) & 7;
            }
            cd.required = req;
            cd.num_required = num & 0xFF;
            cd.t0 = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s, node, 0)[0]);
            cd.w0 = clamp01(sp_vals(s, node, 1)[0]);
            cd.e0 = fmaxf(sp_vals(s, node, 2)[0], 0.f);
        }
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
        const int node = cd.node;
        for (int turn = 0; turn < T; turn++) {
            const float tp = cur_shanten == 0 ? 1.f : clamp01(sp_vals(s, node, 0)[turn]);
            if (!(tp > 0.f)) break;
            const float wp = clamp01(sp_vals(s, node, 1)[turn]);
            const float ev = fminf(SP_FMUL(fmaxf(sp_vals(s, node, 2)[turn], 0.f), ev_scale), 1.f);
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
