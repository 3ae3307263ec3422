// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi state/obs_repr.rs:27-774 (+ array.rs Simple2DArray semantics).
#include "board.h"
#include "sp.h"

#include <cmath>

namespace orc {

// consts.rs:20-28
int obs_rows(int version) {
    switch (version) {
        case 1: return 938;
        case 2: return 942;
        case 3: return 934;
        case 4: return 1012;
        default: throw OrcError("bad obs version");
    }
}

namespace {

const int SELF_KAWA_ITEM_CHANNELS = 4;
const int KAWA_ITEM_CHANNELS = 8;
const int MAX_NUM_TURNS = 17;

struct Ctx {
    const PlayerState& st;
    float* arr;
    int rows;
    int idx = 0;
    bool at_kan_select;
    int version;

    // array.rs:1-64
    void fill(int row, float v) { check(row); for (int c = 0; c < 34; c++) arr[row * 34 + c] = v; }
    void assign(int row, int col, float v) { check(row); arr[row * 34 + col] = v; }
    float get(int row, int col) { check(row); return arr[row * 34 + col]; }
    void fill_rows(int start, int n, float v) { for (int i = 0; i < n; i++) fill(start + i, v); }
    void assign_rows(int start, int col, int n, float v) { for (int i = 0; i < n; i++) assign(start + i, col, v); }
    void check(int row) { if (row < 0 || row >= rows) throw OrcError("obs row out of range"); }

    // obs_repr.rs:27-108 IntegerEncoder
    void int_enc(int n_raw, int cap, bool one_hot, bool rescale, int rbf_intervals) {
        int n = std::min(n_raw, cap);
        switch (version) {
            case 1:
                fill_rows(idx, n, 1.f);
                idx += cap;
                break;
            case 2: case 3:
                if (one_hot) { fill(idx + n, 1.f); idx += cap + 1; }
                if (rescale) { fill(idx, (float)n / (float)cap); idx += 1; }
                if (rbf_intervals > 0) {
                    float interval_size = (float)cap / (float)rbf_intervals;
                    for (int i = 1; i < rbf_intervals; i++) {
                        float x = (float)n_raw;
                        float mu = (float)i * interval_size;
                        float sigma = interval_size;
                        float d = x - mu;
                        float v = std::exp(-(d * d) / (2.f * (sigma * sigma)));
                        fill(idx + i - 1, v);
                    }
                    idx += rbf_intervals - 1;
                }
                break;
            case 4:
                if (one_hot) { fill(idx + n, 1.f); idx += cap + 1; }
                if (rescale) { fill(idx, (float)n / (float)cap); idx += 1; }
                break;
        }
    }

    // obs_repr.rs:694-712
    void encode_tile_set(const std::vector<u8>& tiles) {
        int counts[34] = {0};
        for (u8 tile : tiles) {
            int tid = deaka(tile);
            assign(idx + counts[tid], tid, 1.f);
            counts[tid]++;
            if (is_aka(tile)) fill(idx + 4 + (tile - T_5MR), 1.f);
        }
        idx += 7;
    }

    // obs_repr.rs:714-734
    void encode_self_kawa(const KawaSlot& k) {
        if (k.some) {
            for (int i = 0; i < k.item.n_kan; i++) assign(idx, deaka(k.item.kan[i]), 1.f);
            const Sutehai& s = k.item.sutehai;
            assign(idx + 1, deaka(s.tile), 1.f);
            if (is_aka(s.tile)) fill(idx + 2, 1.f);
            if (s.is_dora) fill(idx + 3, 1.f);
        }
        idx += SELF_KAWA_ITEM_CHANNELS;
    }

    // obs_repr.rs:736-774
    void encode_kawa(const KawaSlot& k) {
        if (k.some) {
            if (k.item.has_chi_pon) {
                int a = deaka(k.item.chi_pon.consumed[0]), b = deaka(k.item.chi_pon.consumed[1]);
                assign(idx, std::min(a, b), 1.f);
                assign(idx + 1, std::max(a, b), 1.f);
            }
            for (int i = 0; i < k.item.n_kan; i++) assign(idx + 2, deaka(k.item.kan[i]), 1.f);
            const Sutehai& s = k.item.sutehai;
            assign(idx + 3, deaka(s.tile), 1.f);
            if (is_aka(s.tile)) fill(idx + 4, 1.f);
            if (s.is_dora) fill(idx + 5, 1.f);
            if (s.is_tedashi) fill(idx + 6, 1.f);
            if (s.is_riichi) fill(idx + 7, 1.f);
        }
        idx += KAWA_ITEM_CHANNELS;
    }

    // obs_repr.rs:632-638
    void encode_ev(float value) {
        fill(idx, std::min(std::max(value, 0.f), 100000.f) / 100000.f);
        fill(idx + 1, std::min(std::max(value, 0.f), 30000.f) / 30000.f);
        idx += 2;
    }

    // obs_repr.rs:644-692
    void encode_sp_table(const std::vector<SpCandidate>& cands, bool can_discard, float ev_scale) {
        bool ok = !cands.empty() && !cands[0].tenpai_probs.empty() && cands[0].tenpai_probs[0] > 0.f;
        if (!ok) { idx += 3 * MAX_NUM_TURNS; return; }
        auto emit = [&](const SpCandidate& c, int tid_or_fill) {
            size_t n = std::min(std::min(c.tenpai_probs.size(), c.win_probs.size()), c.exp_values.size());
            for (size_t turn = 0; turn < n; turn++) {
                if (!(c.tenpai_probs[turn] > 0.f)) break;
                int i = idx + (int)turn;
                float ev = std::min(c.exp_values[turn] * ev_scale, 1.f);
                if (tid_or_fill >= 0) {
                    assign(i, tid_or_fill, c.tenpai_probs[turn]);
                    assign(i + MAX_NUM_TURNS, tid_or_fill, c.win_probs[turn]);
                    assign(i + 2 * MAX_NUM_TURNS, tid_or_fill, ev);
                } else {
                    fill(i, c.tenpai_probs[turn]);
                    fill(i + MAX_NUM_TURNS, c.win_probs[turn]);
                    fill(i + 2 * MAX_NUM_TURNS, ev);
                }
            }
        };
        if (can_discard) for (auto& c : cands) emit(c, deaka(c.tile));
        else emit(cands[0], -1);
        idx += 3 * MAX_NUM_TURNS;
    }
};

}  // namespace

// obs_repr.rs:126-630
void PlayerState::encode_obs(int version, bool at_kan_select, float* obs, u8* mask, int sp_mode) const {
    const PlayerState& state = *this;
    int rows = obs_rows(version);
    memset(obs, 0, sizeof(float) * rows * 34);
    memset(mask, 0, 46);
    Ctx c{state, obs, rows, 0, at_kan_select, version};
    const ActionCandidate& cans = last_cans;

    for (int t = 0; t < 34; t++) if (tehai[t] > 0) c.assign_rows(c.idx, t, tehai[t], 1.f);
    c.idx += 4;
    for (int i = 0; i < 3; i++) if (akas_in_hand[i]) c.fill(c.idx + i, 1.f);
    c.idx += 3;

    for (int i = 0; i < 4; i++) {
        i32 score = scores[i];
        c.fill(c.idx, (float)std::min(std::max(score, 0), 100000) / 100000.f);
        c.idx += 1;
        if (version == 2 || version == 3) {
            // `score as usize / 100`: negative scores wrap in Rust; the reference only feeds
            // non-negative scores here in self-play (tobi ends the game first).
            c.int_enc((int)((u32)score / 100), 500, false, false, 10);
        } else if (version == 4) {
            c.fill(c.idx, (float)std::min(std::max(score, 0), 30000) / 30000.f);
            c.idx += 1;
        }
    }

    c.fill(c.idx + rank, 1.f);
    c.idx += 4;

    if (version == 1) c.fill_rows(c.idx, kyoku, 1.f);
    else c.fill(c.idx + kyoku, 1.f);
    c.idx += 4;

    int cap = (version == 1 || version == 4) ? 10 : 6;
    c.int_enc(honba, cap, false, version == 4, 3);
    c.int_enc(kyotaku, cap, false, version == 4, 3);

    c.assign(c.idx, bakaze, 1.f);
    c.assign(c.idx + 1, jikaze, 1.f);
    c.idx += 2;

    if (version >= 2) {
        int n = std::min<int>(bakaze - T_E, 1) * 4 + kyoku;
        c.int_enc(n, 7, false, true, 0);
    }

    c.encode_tile_set(dora_indicators);

    {
        const auto& k0 = kawa[0];
        size_t n = std::min<size_t>(k0.size(), 6);
        for (size_t i = 0; i < n; i++) c.encode_self_kawa(k0[i]);
        c.idx += (int)(6 - n) * SELF_KAWA_ITEM_CHANNELS;
        n = std::min<size_t>(k0.size(), 18);
        for (size_t i = 0; i < n; i++) c.encode_self_kawa(k0[k0.size() - 1 - i]);
        c.idx += (int)(18 - n) * SELF_KAWA_ITEM_CHANNELS;
    }

    size_t max_kawa_len = 0;
    for (int i = 0; i < 4; i++) max_kawa_len = std::max(max_kawa_len, kawa[i].size());
    if (version >= 3) {
        for (size_t turn = 0; turn < kawa[0].size(); turn++) {
            if (!kawa[0][turn].some) continue;
            int tid = deaka(kawa[0][turn].item.sutehai.tile);
            float v = std::exp(-0.2f * (float)(max_kawa_len - 1 - turn));
            c.assign(c.idx, tid, v);
        }
        c.idx += 1;
    }

    for (int p = 1; p < 4; p++) {
        const auto& pk = kawa[p];
        size_t n = std::min<size_t>(pk.size(), 6);
        for (size_t i = 0; i < n; i++) c.encode_kawa(pk[i]);
        c.idx += (int)(6 - n) * KAWA_ITEM_CHANNELS;
        n = std::min<size_t>(pk.size(), 18);
        for (size_t i = 0; i < n; i++) c.encode_kawa(pk[pk.size() - 1 - i]);
        c.idx += (int)(18 - n) * KAWA_ITEM_CHANNELS;

        if (version == 2) {
            int turn = 0;
            for (const auto& slot : pk) {
                if (!slot.some) continue;
                int row = std::min(turn / 6, 2);
                int tid = deaka(slot.item.sutehai.tile);
                c.assign(c.idx + row, tid, 1.f);
                if (slot.item.sutehai.is_tedashi) c.assign(c.idx + 3 + row, tid, 1.f);
                turn++;
            }
            c.idx += 6;
        } else if (version >= 3) {
            for (size_t turn = 0; turn < pk.size(); turn++) {
                if (!pk[turn].some) continue;
                const Sutehai& s = pk[turn].item.sutehai;
                int tid = deaka(s.tile);
                float v = std::exp(-0.2f * (float)(max_kawa_len - 1 - turn));
                c.assign(c.idx, tid, v);
                if (s.is_tedashi) c.assign(c.idx + 1, tid, v);
                if (s.is_riichi) c.assign(c.idx + 2, tid, v);
            }
            c.idx += 3;
        }
    }

    c.fill(c.idx, (float)tiles_left / 69.f);
    c.idx += 1;

    for (int i = 0; i < 4; i++) c.int_enc(doras_owned[i], 12, false, true, 3);

    u8 doras_unseen = (u8)(dora_indicators.size() * 4 + 3 - doras_seen);
    c.int_enc(doras_unseen, 5 * 4 + 3, false, true, 4);

    for (int p = 0; p < 4; p++) c.encode_tile_set(kawa_overview[p]);

    for (int p = 0; p < 4; p++) {
        for (const auto& f : fuuro_overview[p]) {
            for (u8 tile : f) {
                int tid = deaka(tile);
                int i = 0;
                while (i < 4 && c.get(c.idx + i, tid) != 0.f) i++;
                if (i >= 4) throw OrcError("fuuro encode: no free row");
                c.assign(c.idx + i, tid, 1.f);
                if (is_aka(tile)) c.fill(c.idx + 4, 1.f);
            }
            c.idx += 5;
        }
        c.idx += (int)(4 - fuuro_overview[p].size()) * 5;
    }

    for (int p = 0; p < 4; p++) {
        for (u8 tile : ankan_overview[p]) c.assign(c.idx, tile, 1.f);
        c.idx += 1;
    }

    if (version >= 2) {
        for (int t = 0; t < 34; t++) c.assign(c.idx, t, (float)tiles_seen[t] / 4.f);
        c.idx += 1;
        for (int p = 1; p < 4; p++) {
            if (last_tedashis[p].some) {
                const Sutehai& s = last_tedashis[p].s;
                c.assign(c.idx, deaka(s.tile), 1.f);
                if (is_aka(s.tile)) c.fill(c.idx + 1, 1.f);
                if (s.is_dora) c.fill(c.idx + 2, 1.f);
            }
            c.idx += 3;
        }
        for (int p = 1; p < 4; p++) {
            if (riichi_sutehais[p].some) {
                const Sutehai& s = riichi_sutehais[p].s;
                c.assign(c.idx, deaka(s.tile), 1.f);
                if (is_aka(s.tile)) c.fill(c.idx + 1, 1.f);
                if (s.is_dora) c.fill(c.idx + 2, 1.f);
            }
            c.idx += 3;
        }
    }

    for (int p = 1; p < 4; p++) if (riichi_declared[p]) c.fill(c.idx + p - 1, 1.f);
    c.idx += 3;
    for (int p = 1; p < 4; p++) if (riichi_accepted[p]) c.fill(c.idx + p - 1, 1.f);
    c.idx += 3;

    for (int t = 0; t < 34; t++) if (waits[t]) c.assign(c.idx, t, 1.f);
    c.idx += 1;

    if (at_furiten) c.fill(c.idx, 1.f);
    c.idx += 1;

    c.int_enc(shanten, 6, true, false, 0);

    if (riichi_accepted[0]) c.fill(c.idx, 1.f);
    c.idx += 1;
    if (at_kan_select) c.fill(c.idx, 1.f);
    c.idx += 1;

    if (cans.can_pass()) {
        ORC_ENSURE(has_last_kawa_tile, "building chi/pon/daiminkan/ron feature without any kawa tile");
        u8 tile = last_kawa_tile;
        int tid = deaka(tile);
        c.assign(c.idx, tid, 1.f);
        if (is_aka(tile)) c.fill(c.idx + 1, 1.f);
        if (dora_factor[tid] > 0) c.fill(c.idx + 2, 1.f);
        if (!at_kan_select) mask[45] = 1;
        else if (cans.can_daiminkan) mask[tid] = 1;
    }
    c.idx += 3;

    if (cans.can_discard) {
        bool dc[37];
        discard_candidates_aka(dc);
        for (int t = 0; t < 37; t++) {
            if (!dc[t]) continue;
            c.assign(c.idx, deaka((u8)t), 1.f);
            if (!at_kan_select) mask[t] = 1;
        }
        for (int t = 0; t < 34; t++) if (keep_shanten_discards[t]) c.assign(c.idx + 1, t, 1.f);
        for (int t = 0; t < 34; t++) if (next_shanten_discards[t]) c.assign(c.idx + 2, t, 1.f);
        if (shanten <= 1) {
            bool ut[34];
            discard_candidates_with_unconditional_tenpai(ut);
            for (int t = 0; t < 34; t++) if (ut[t]) c.assign(c.idx + 3, t, 1.f);
        }
        if (riichi_declared[0]) c.fill(c.idx + 4, 1.f);
    }
    c.idx += 5;

    if (cans.can_riichi) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[37] = 1; }
    c.idx += 1;
    if (cans.can_chi_low) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[38] = 1; }
    if (cans.can_chi_mid) { c.fill(c.idx + 1, 1.f); if (!at_kan_select) mask[39] = 1; }
    if (cans.can_chi_high) { c.fill(c.idx + 2, 1.f); if (!at_kan_select) mask[40] = 1; }
    c.idx += 3;
    if (cans.can_pon) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[41] = 1; }
    c.idx += 1;
    if (cans.can_daiminkan) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[42] = 1; }
    c.idx += 1;
    if (cans.can_ankan) {
        for (u8 t : ankan_candidates) { c.assign(c.idx, t, 1.f); if (at_kan_select) mask[t] = 1; }
        if (!at_kan_select) mask[42] = 1;
    }
    c.idx += 1;
    if (cans.can_kakan) {
        for (u8 t : kakan_candidates) { c.assign(c.idx, t, 1.f); if (at_kan_select) mask[t] = 1; }
        if (!at_kan_select) mask[42] = 1;
    }
    c.idx += 1;
    if (cans.can_agari()) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[43] = 1; }
    c.idx += 1;
    if (cans.can_ryukyoku) { c.fill(c.idx, 1.f); if (!at_kan_select) mask[44] = 1; }
    c.idx += 1;

    if (version == 4) {
        std::vector<SpCandidate> table;
        bool have = false;
        if (sp_mode == 1) have = single_player_tables(state, table);
        if (have) {
            float max_ev = (!table.empty() && !table[0].exp_values.empty()) ? table[0].exp_values[0] : 0.f;
            c.encode_ev(max_ev);
            if (cans.can_discard) {
                for (const auto& cand : table) {
                    int dt = deaka(cand.tile);
                    for (const auto& r : cand.required_tiles) {
                        int rt = deaka(r.tile);
                        if (cand.shanten_down) c.assign(c.idx + 34 + dt, rt, 1.f);
                        else c.assign(c.idx + dt, rt, 1.f);
                    }
                }
                c.idx += 2 * 34;
                // max_by(NotShantenDown): last maximum
                int best = 0;
                for (size_t i = 1; i < table.size(); i++)
                    if (sp_candidate_cmp(table[i], table[best], SPCOL_NOT_SHANTEN_DOWN) >= 0) best = (int)i;
                c.assign(c.idx, deaka(table[best].tile), 1.f);
                c.idx += 2;
            } else {
                c.idx += 2 * 34 + 1;
                for (const auto& r : table[0].required_tiles) c.assign(c.idx, deaka(r.tile), 1.f);
                c.idx += 1;
            }
            float ev_scale = max_ev < 1.f ? 0.f : 1.f / max_ev;
            c.encode_sp_table(table, cans.can_discard, ev_scale);
        } else {
            float min_tsumo_agari = 0.f;
            if (sp_mode == 1) {
                try {
                    Point p = agari_points(cans.can_ron_agari, nullptr, 0);
                    min_tsumo_agari = (float)p.tsumo_total(is_oya());
                } catch (const OrcError&) {
                    min_tsumo_agari = 0.f;
                }
                c.encode_ev(min_tsumo_agari);
            } else {
                c.idx += 2;  // sp_mode 0: SP block left entirely zero (documented gap mode)
            }
            c.idx += 2 * 34 + 2 + 3 * MAX_NUM_TURNS;
        }
    }

    if (c.idx != rows) throw OrcError("obs encode: row count mismatch");
}

}  // namespace orc
