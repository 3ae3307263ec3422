// TEST INFRASTRUCTURE ONLY — single-lane host build of mortal_b200/csrc/mjx_step.cuh (-DMJX_HOST_EMUL).
// Lets the GPU-less dev container diff the product's rule logic against the oracle. Never loaded by
// mortal_b200/; the product has no CPU path.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "mjx_step.cuh"
#include "mjx_policy.cuh"
#include "mjx_tables_host.h"

using namespace mjx;

static HostTables g_H;
static Tables g_T;
static bool g_ready = false;
static std::string g_err;

extern "C" {

const char* emul_last_error() { return g_err.c_str(); }

int emul_init(const char* data_dir) {
    if (g_ready) return 0;
    if (!load_host_tables(data_dir, g_H)) { g_err = g_H.error; return -1; }
    g_T.suhai = g_H.suhai.data();
    g_T.jihai = g_H.jihai.data();
    g_T.agari_keys = g_H.agari_keys.data();
    g_T.agari_divs = g_H.agari_divs.data();
    g_T.agari_ndivs = g_H.agari_ndivs.data();
    g_ready = true;
    return 0;
}

int emul_sizeof_table() { return (int)sizeof(TableState); }

int emul_shanten(const uint8_t* tiles, const uint8_t* len_div3, int8_t* out, int n) {
    for (int i = 0; i < n; i++) out[i] = (int8_t)shanten_all(g_T, tiles + (size_t)i * 34, len_div3[i]);
    return 0;
}

void emul_make_wall(uint64_t nonce, uint64_t key, int kyoku, int honba, int kind, uint8_t* out) {
    make_wall(nonce, key, kyoku, honba, kind, out);
}

// trace rows: [table, step_idx, seat, action, kan_select, mask_bits]
int emul_run(int n, const uint64_t* nonces, const uint64_t* keys, int shuffle_kind, int quick_eval, int policy_kind,
             int32_t* scores, uint8_t* ranks, int32_t* steps, int32_t* errs, int64_t* trace, int64_t trace_cap,
             int64_t* trace_len, int64_t max_cycles) {
    std::vector<TableState> tabs(n);
    for (int t = 0; t < n; t++) {
        TableState& S = tabs[t];
        memset(&S, 0, sizeof S);
        S.nonce = nonces[t]; S.key = keys[t];
        for (int i = 0; i < 4; i++) { S.scores[i] = 25000; S.row_of_seat[i] = -1; S.kan_row_of_seat[i] = -1; S.auto_action[i] = -1; }
        S.shuffle_kind = (u8)shuffle_kind;
        S.gflags = GF_ALIVE;
    }
    const int cap = n * MAX_ROWS_PER_TABLE;
    std::vector<i32> row_table(cap), done(n, 0), n_rows(1, 0);
    std::vector<u8> row_seat(cap), masks((size_t)cap * ACTION_SPACE);
    std::vector<u32> row_step(cap);
    std::vector<i64> actions(cap, 45);
    unsigned long long counters[2] = {0, 0};
    EnvView V;
    V.tables = tabs.data(); V.n_tables = n; V.row_cap = cap; V.n_rows = n_rows.data();
    V.row_table = row_table.data(); V.row_seat = row_seat.data(); V.row_step = row_step.data();
    V.masks = masks.data(); V.actions = actions.data(); V.scores = scores; V.ranks = ranks; V.done = done.data();
    V.steps = steps; V.err = errs; V.counters = counters; V.enable_quick_eval = quick_eval;
    for (int t = 0; t < n; t++) { steps[t] = 0; errs[t] = 0; }
    int64_t tl = 0;
    WarpScratch W;
    for (int64_t cycle = 0; max_cycles <= 0 || cycle < max_cycles; cycle++) {
        n_rows[0] = 0;
        int live = 0;
        for (int t = 0; t < n; t++) {
            Ctx c;
            c.S = &tabs[t]; c.W = &W; c.T = g_T; c.lane = 0;
            if (step_table(c, V, t)) live++;
        }
        if (live == 0) break;
        // policy over the emitted rows
        for (int r = 0; r < n_rows[0]; r++) {
            int t = row_table[r], seat = row_seat[r] & 3, kan = (row_seat[r] >> 2) & 1;
            u64 m = 0;
            for (int i = 0; i < ACTION_SPACE; i++) if (masks[(size_t)r * ACTION_SPACE + i]) m |= 1ull << i;
            const SeatPrivate& P = tabs[t].priv[seat];
            u64 h = policy_hash(tabs[t].nonce, tabs[t].key, (u64)t, row_step[r], (u32)seat, (u32)kan);
            int a = test_policy(policy_kind, h, kan != 0, m, P.keep_shanten, P.next_shanten);
            actions[r] = a;
            if (trace && tl < trace_cap) {
                int64_t* o = trace + tl * 6;
                o[0] = t; o[1] = row_step[r]; o[2] = seat; o[3] = a; o[4] = kan; o[5] = (int64_t)m;
            }
            tl++;
        }
    }
    if (trace_len) *trace_len = tl;
    for (int t = 0; t < n; t++) if (!done[t]) { errs[t] = errs[t] ? errs[t] : -1; }
    return 0;
}

}  // extern "C"
