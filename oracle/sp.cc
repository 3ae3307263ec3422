// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi algo/sp/{calc,state,candidate}.rs (default features, i.e. not
// `sp_reproduce_cpp_ver`) and state/agent_helper.rs:509-593.
// f32 evaluation order follows the Rust source statement by statement; build with
// -ffp-contract=off so no FMA contraction changes the rounding.
#include "sp.h"

#include <algorithm>
#include <memory>
#include <unordered_map>

namespace orc {

long g_sp_stats[8] = {0,0,0,0,0,0,0,0};  // calls, sum states, max states, calls>1k, >10k, >100k

namespace {

const int SHANTEN_THRES = 3;                   // calc.rs:13
const int MAX_TILES_LEFT = 34 * 4 - 1 - 13;    // calc.rs:14
const int MAX_TSUMOS_LEFT = 17;                // sp/mod.rs:42

// data/uradora_prob_table.txt (calc.rs:17)
const float URADORA_PROB_TABLE[5][13] = {
    {0.639485f, 0.327801f, 0.0327134f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.406736f, 0.42281f, 0.147966f, 0.021674f, 0.0008142f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.257516f, 0.406819f, 0.246851f, 0.0757724f, 0.0122266f, 0.0008004f, 1.43e-5f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.162199f, 0.346513f, 0.301539f, 0.142396f, 0.0401276f, 0.0066491f, 0.0005575f, 1.85e-5f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.101768f, 0.275319f, 0.313742f, 0.20189f, 0.081774f, 0.0215394f, 0.0035918f, 0.0003607f, 1.52e-5f, 3e-7f, 0.f, 0.f, 0.f},
};

// sp/state.rs:10-21
struct State {
    u8 tehai[34];
    u8 akas_in_hand[3];
    u8 tiles_in_wall[34];
    u8 akas_in_wall[3];
    u8 n_extra_tsumo;
    bool operator==(const State& o) const { return memcmp(this, &o, sizeof(State)) == 0; }
};
struct StateHash {
    size_t operator()(const State& s) const {
        const u8* p = reinterpret_cast<const u8*>(&s);
        u64 h = 1469598103934665603ULL;
        for (size_t i = 0; i < sizeof(State); i++) { h ^= p[i]; h *= 1099511628211ULL; }
        return (size_t)h;
    }
};

struct DiscardTile { u8 tile; i8 shanten_diff; };
struct DrawTile { u8 tile; u8 count; i8 shanten_diff; };

// sp/state.rs:57-104
void st_discard(State& s, u8 tile) {
    s.tehai[deaka(tile)] -= 1;
    if (is_aka(tile)) s.akas_in_hand[tile - T_5MR] = 0;
}
void st_undo_discard(State& s, u8 tile) {
    s.tehai[deaka(tile)] += 1;
    if (is_aka(tile)) s.akas_in_hand[tile - T_5MR] = 1;
}
void st_deal(State& s, u8 tile) {
    s.tiles_in_wall[deaka(tile)] -= 1;
    if (is_aka(tile)) s.akas_in_wall[tile - T_5MR] = 0;
    st_undo_discard(s, tile);
}
void st_undo_deal(State& s, u8 tile) {
    st_discard(s, tile);
    s.tiles_in_wall[deaka(tile)] += 1;
    if (is_aka(tile)) s.akas_in_wall[tile - T_5MR] = 1;
}

// sp/state.rs:106-136
std::vector<DiscardTile> get_discard_tiles(const State& s, i8 shanten, u8 len_div3) {
    std::vector<DiscardTile> out;
    u8 tehai[34];
    memcpy(tehai, s.tehai, 34);
    for (int tid = 0; tid < 34; tid++) {
        if (tehai[tid] == 0) continue;
        tehai[tid] -= 1;
        i8 after = shanten_all(tehai, len_div3);
        tehai[tid] += 1;
        i8 diff = after - shanten;
        u8 tile = (u8)tid;
        if (tid == T_5M && s.akas_in_hand[0] && tehai[tid] == 1) tile = T_5MR;
        else if (tid == T_5P && s.akas_in_hand[1] && tehai[tid] == 1) tile = T_5PR;
        else if (tid == T_5S && s.akas_in_hand[2] && tehai[tid] == 1) tile = T_5SR;
        out.push_back({tile, diff});
    }
    return out;
}

// sp/state.rs:138-179
std::vector<DrawTile> get_draw_tiles(const State& s, i8 shanten, u8 len_div3) {
    std::vector<DrawTile> out;
    u8 tehai[34];
    memcpy(tehai, s.tehai, 34);
    for (int tid = 0; tid < 34; tid++) {
        u8 count = s.tiles_in_wall[tid];
        if (count == 0) continue;
        tehai[tid] += 1;
        i8 after = shanten_all(tehai, len_div3);
        tehai[tid] -= 1;
        i8 diff = after - shanten;
        bool aka_in_wall = (tid == T_5M && s.akas_in_wall[0]) || (tid == T_5P && s.akas_in_wall[1]) ||
                           (tid == T_5S && s.akas_in_wall[2]);
        if (aka_in_wall) {
            if (count >= 2) out.push_back({(u8)tid, (u8)(count - 1), diff});
            out.push_back({akaize((u8)tid), 1, diff});
        } else {
            out.push_back({(u8)tid, count, diff});
        }
    }
    return out;
}

// sp/state.rs:181-201
std::vector<RequiredTile> get_required_tiles(const State& s, u8 len_div3) {
    u8 tehai[34];
    memcpy(tehai, s.tehai, 34);
    i8 shanten = shanten_all(tehai, len_div3);
    std::vector<RequiredTile> out;
    for (int tid = 0; tid < 34; tid++) {
        u8 count = s.tiles_in_wall[tid];
        if (count == 0) continue;
        tehai[tid] += 1;
        i8 after = shanten_all(tehai, len_div3);
        tehai[tid] -= 1;
        if (after < shanten) out.push_back({(u8)tid, count});
    }
    return out;
}

u8 sum_left_tiles(const State& s) {
    u8 n = 0;
    for (int i = 0; i < 34; i++) n += s.tiles_in_wall[i];
    return n;
}

struct Values {
    std::vector<float> tenpai_probs, win_probs, exp_values;
    explicit Values(int T) : tenpai_probs(T, 0.f), win_probs(T, 0.f), exp_values(T, 0.f) {}
};
typedef std::shared_ptr<Values> ValuesPtr;
typedef std::unordered_map<State, ValuesPtr, StateHash> Cache;

// sp/candidate.rs:49-71
SpCandidate make_candidate(u8 tile, const std::vector<float>* tp, const std::vector<float>* wp,
                           const std::vector<float>* ev, std::vector<RequiredTile> req, bool shanten_down) {
    SpCandidate c;
    c.tile = tile;
    if (tp) for (float p : *tp) c.tenpai_probs.push_back(std::min(std::max(p, 0.f), 1.f));
    if (wp) for (float p : *wp) c.win_probs.push_back(std::min(std::max(p, 0.f), 1.f));
    if (ev) for (float v : *ev) c.exp_values.push_back(std::max(v, 0.f));
    u8 n = 0;
    for (auto& r : req) n += r.count;
    c.num_required_tiles = n;
    c.required_tiles = std::move(req);
    c.shanten_down = shanten_down;
    return c;
}

struct CalcState {
    const SpCalculator& sup;
    State state;
    int T;  // MAX_TSUMO
    std::vector<std::vector<float>> tsumo_prob_table;      // [4][T]
    std::vector<std::vector<float>> not_tsumo_prob_table;  // [MAX_TILES_LEFT+1][T]
    Cache discard_cache[SHANTEN_THRES + 1], draw_cache[SHANTEN_THRES + 1];

    CalcState(const SpCalculator& s, const State& st, int T_, int n_left) : sup(s), state(st), T(T_) {
        // calc.rs:136-146
        tsumo_prob_table.assign(4, std::vector<float>(T, 0.f));
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < T; j++) tsumo_prob_table[i][j] = (float)(i + 1) / (float)(n_left - j);
        // calc.rs:148-167
        not_tsumo_prob_table.assign(MAX_TILES_LEFT + 1, std::vector<float>(T, 0.f));
        for (int i = 0; i <= n_left && i <= MAX_TILES_LEFT; i++) {
            auto& row = not_tsumo_prob_table[i];
            row[0] = 1.f;
            int lim = std::min(T - 1, n_left - i);
            for (int j = 0; j < lim; j++) row[j + 1] = row[j] * (float)(n_left - i - j) / (float)(n_left - j);
        }
    }

    // calc.rs:169-201
    std::vector<SpCandidate> calc(bool can_discard, i8 cur_shanten) {
        std::vector<SpCandidate> cands;
        if (cur_shanten <= SHANTEN_THRES) {
            cands = can_discard ? analyze_discard(cur_shanten) : analyze_draw(cur_shanten);
            if (sup.sort_result && !cands.empty()) {
                SpColumn by = sup.maximize_win_prob ? SPCOL_WIN_PROB : SPCOL_EV;
                std::stable_sort(cands.begin(), cands.end(),
                                 [&](const SpCandidate& l, const SpCandidate& r) { return sp_candidate_cmp(r, l, by) < 0; });
            }
        } else {
            cands = can_discard ? analyze_discard_simple(cur_shanten) : analyze_draw_simple();
            if (sup.sort_result && !cands.empty()) {
                std::stable_sort(cands.begin(), cands.end(), [&](const SpCandidate& l, const SpCandidate& r) {
                    return sp_candidate_cmp(r, l, SPCOL_NOT_SHANTEN_DOWN) < 0;
                });
            }
        }
        return cands;
    }

    // calc.rs:203-253
    std::vector<SpCandidate> analyze_discard(i8 shanten) {
        auto discard_tiles = get_discard_tiles(state, shanten, sup.tehai_len_div3);
        std::vector<SpCandidate> cands;
        for (auto& dt : discard_tiles) {
            if (dt.shanten_diff == 0) {
                st_discard(state, dt.tile);
                auto req = get_required_tiles(state, sup.tehai_len_div3);
                ValuesPtr v = draw(shanten);
                st_undo_discard(state, dt.tile);
                std::vector<float> tp = v->tenpai_probs;
                if (shanten == 0) std::fill(tp.begin(), tp.end(), 1.f);
                cands.push_back(make_candidate(dt.tile, &tp, &v->win_probs, &v->exp_values, req, false));
            } else if (sup.calc_shanten_down && dt.shanten_diff == 1 && shanten < SHANTEN_THRES) {
                st_discard(state, dt.tile);
                auto req = get_required_tiles(state, sup.tehai_len_div3);
                state.n_extra_tsumo += 1;
                ValuesPtr v = draw(shanten + 1);
                state.n_extra_tsumo -= 1;
                st_undo_discard(state, dt.tile);
                cands.push_back(make_candidate(dt.tile, &v->tenpai_probs, &v->win_probs, &v->exp_values, req, true));
            }
        }
        return cands;
    }

    // calc.rs:255-279
    std::vector<SpCandidate> analyze_draw(i8 shanten) {
        auto req = get_required_tiles(state, sup.tehai_len_div3);
        ValuesPtr v = draw(shanten);
        std::vector<float> tp = v->tenpai_probs;
        if (shanten == 0) std::fill(tp.begin(), tp.end(), 1.f);
        return {make_candidate(T_UNK, &tp, &v->win_probs, &v->exp_values, req, false)};
    }

    // calc.rs:281-303
    std::vector<SpCandidate> analyze_discard_simple(i8 shanten) {
        auto discard_tiles = get_discard_tiles(state, shanten, sup.tehai_len_div3);
        std::vector<SpCandidate> cands;
        for (auto& dt : discard_tiles) {
            st_discard(state, dt.tile);
            auto req = get_required_tiles(state, sup.tehai_len_div3);
            st_undo_discard(state, dt.tile);
            cands.push_back(make_candidate(dt.tile, nullptr, nullptr, nullptr, req, dt.shanten_diff == 1));
        }
        return cands;
    }

    // calc.rs:305-314
    std::vector<SpCandidate> analyze_draw_simple() {
        auto req = get_required_tiles(state, sup.tehai_len_div3);
        return {make_candidate(T_UNK, nullptr, nullptr, nullptr, req, false)};
    }

    // calc.rs:316-322
    ValuesPtr draw(i8 shanten) {
        if (sup.calc_tegawari && state.n_extra_tsumo == 0) return draw_with_tegawari(shanten);
        return draw_without_tegawari(shanten);
    }

    // calc.rs:324-445
    ValuesPtr draw_with_tegawari(i8 shanten) {
        auto it = draw_cache[shanten].find(state);
        if (it != draw_cache[shanten].end()) return it->second;

        ValuesPtr out = std::make_shared<Values>(T);
        auto draw_tiles = get_draw_tiles(state, shanten, sup.tehai_len_div3);
        u8 sum_left = sum_left_tiles(state);

        for (auto& d : draw_tiles) {
            if (d.shanten_diff != -1) continue;
            st_deal(state, d.tile);
            bool is_scores = false;
            float scores[4] = {0, 0, 0, 0};
            ValuesPtr next;
            if (shanten > 0) {
                next = discard(shanten - 1);
            } else if (get_score(d.tile, scores)) {
                is_scores = true;
            } else {
                st_undo_deal(state, d.tile);
                continue;
            }
            st_undo_deal(state, d.tile);

            for (int i = 0; i < T; i++) {
                float tump_prob = (float)d.count / (float)sum_left;
                if (is_scores) {
                    bool assume_riichi = sup.is_menzen && sup.prefer_riichi;
                    bool win_double_riichi = assume_riichi && sup.calc_double_riichi && i == 0;
                    bool win_ippatsu = assume_riichi;
                    bool win_haitei = sup.calc_haitei && i == T - 1;
                    int han_plus = (int)win_double_riichi + (int)win_ippatsu + (int)win_haitei;
                    out->win_probs[i] += tump_prob;
                    out->exp_values[i] += tump_prob * scores[han_plus];
                } else {
                    if (shanten == 1) out->tenpai_probs[i] += tump_prob;
                    if (i < T - 1) {
                        if (shanten > 1) out->tenpai_probs[i] += tump_prob * next->tenpai_probs[i + 1];
                        out->win_probs[i] += tump_prob * next->win_probs[i + 1];
                        out->exp_values[i] += tump_prob * next->exp_values[i + 1];
                    }
                }
            }
        }

        for (auto& d : draw_tiles) {
            if (d.shanten_diff != 0) continue;
            st_deal(state, d.tile);
            state.n_extra_tsumo += 1;
            ValuesPtr next = discard(shanten);
            state.n_extra_tsumo -= 1;
            st_undo_deal(state, d.tile);
            for (int i = 0; i < T - 1; i++) {
                float tump_prob = (float)d.count / (float)sum_left;
                out->tenpai_probs[i] += tump_prob * next->tenpai_probs[i + 1];
                out->win_probs[i] += tump_prob * next->win_probs[i + 1];
                out->exp_values[i] += tump_prob * next->exp_values[i + 1];
            }
        }
        draw_cache[shanten][state] = out;
        return out;
    }

    // calc.rs:447-561
    ValuesPtr draw_without_tegawari(i8 shanten) {
        auto it = draw_cache[shanten].find(state);
        if (it != draw_cache[shanten].end()) return it->second;

        ValuesPtr out = std::make_shared<Values>(T);
        auto draw_tiles = get_draw_tiles(state, shanten, sup.tehai_len_div3);
        u8 sum_required = 0;
        for (auto& d : draw_tiles) if (d.shanten_diff == -1) sum_required += d.count;
        const std::vector<float>& not_tsumo_probs = not_tsumo_prob_table[sum_required];

        for (auto& d : draw_tiles) {
            if (d.shanten_diff != -1) continue;
            st_deal(state, d.tile);
            bool is_scores = false;
            float scores[4] = {0, 0, 0, 0};
            ValuesPtr next;
            if (shanten > 0) {
                next = discard(shanten - 1);
            } else if (get_score(d.tile, scores)) {
                is_scores = true;
            } else {
                st_undo_deal(state, d.tile);
                continue;
            }
            st_undo_deal(state, d.tile);

            const std::vector<float>& tsumo_probs = tsumo_prob_table[d.count - 1];
            for (int i = 0; i < T; i++) {
                float m = not_tsumo_probs[i];
                if (m == 0.f) break;
                for (int j = i; j < T; j++) {
                    float n = not_tsumo_probs[j];
                    if (n == 0.f) break;
                    float prob = tsumo_probs[j] * n / m;
                    if (is_scores) {
                        bool assume_riichi = sup.is_menzen && sup.prefer_riichi;
                        bool win_double_riichi = assume_riichi && sup.calc_double_riichi && i == 0;
                        bool win_ippatsu = assume_riichi && j == i;
                        bool win_haitei = sup.calc_haitei && j == T - 1;
                        int han_plus = (int)win_double_riichi + (int)win_ippatsu + (int)win_haitei;
                        out->win_probs[i] += prob;
                        out->exp_values[i] += prob * scores[han_plus];
                    } else {
                        if (shanten == 1) out->tenpai_probs[i] += prob;
                        if (j < T - 1) {
                            if (shanten > 1) out->tenpai_probs[i] += prob * next->tenpai_probs[j + 1];
                            out->win_probs[i] += prob * next->win_probs[j + 1];
                            out->exp_values[i] += prob * next->exp_values[j + 1];
                        }
                    }
                }
            }
        }
        draw_cache[shanten][state] = out;
        return out;
    }

    // calc.rs:563-637
    ValuesPtr discard(i8 shanten) {
        auto it = discard_cache[shanten].find(state);
        if (it != discard_cache[shanten].end()) return it->second;

        auto discard_tiles = get_discard_tiles(state, shanten, sup.tehai_len_div3);
        const float FMIN = -3.40282347e+38f;  // f32::MIN
        ValuesPtr out = std::make_shared<Values>(T);
        std::fill(out->tenpai_probs.begin(), out->tenpai_probs.end(), FMIN);
        std::fill(out->win_probs.begin(), out->win_probs.end(), FMIN);
        std::fill(out->exp_values.begin(), out->exp_values.end(), FMIN);
        std::vector<u8> max_tiles(T, T_UNK);
        std::vector<i32> max_values(T, INT32_MIN);

        for (auto& dt : discard_tiles) {
            ValuesPtr values;
            if (dt.shanten_diff == 0) {
                st_discard(state, dt.tile);
                values = draw(shanten);
                st_undo_discard(state, dt.tile);
            } else if (sup.calc_shanten_down && state.n_extra_tsumo == 0 && dt.shanten_diff == 1 &&
                       shanten < SHANTEN_THRES) {
                st_discard(state, dt.tile);
                state.n_extra_tsumo += 1;
                values = draw(shanten + 1);
                state.n_extra_tsumo -= 1;
                st_undo_discard(state, dt.tile);
            } else {
                continue;
            }
            for (int i = 0; i < T; i++) {
                float fv = sup.maximize_win_prob ? values->win_probs[i] * 1e5f : values->exp_values[i];
                i32 value = (i32)fv;
                if (value > max_values[i] ||
                    (value == max_values[i] && cmp_discard_priority(dt.tile, max_tiles[i]) > 0)) {
                    out->tenpai_probs[i] = values->tenpai_probs[i];
                    out->win_probs[i] = values->win_probs[i];
                    out->exp_values[i] = values->exp_values[i];
                    max_values[i] = value;
                    max_tiles[i] = dt.tile;
                }
            }
        }
        discard_cache[shanten][state] = out;
        return out;
    }

    // calc.rs:640-758
    bool get_score(u8 win_tile, float* scores) {
        AgariCalc calc;
        calc.tehai = state.tehai;
        calc.is_menzen = sup.is_menzen;
        calc.chis = sup.chis; calc.n_chis = sup.n_chis;
        calc.pons = sup.pons; calc.n_pons = sup.n_pons;
        calc.minkans = sup.minkans; calc.n_minkans = sup.n_minkans;
        calc.ankans = sup.ankans; calc.n_ankans = sup.n_ankans;
        calc.bakaze = sup.bakaze; calc.jikaze = sup.jikaze;
        calc.winning_tile = deaka(win_tile);
        calc.is_ron = false;
        bool is_oya = sup.jikaze == T_E;

        u8 additional_yakus = sup.is_menzen ? (sup.prefer_riichi ? 2 : 1) : 0;
        u8 num_doras = 0;
        for (int i = 0; i < sup.n_dora_indicators; i++) num_doras += state.tehai[tile_next(sup.dora_indicators[i])];
        for (int i = 0; i < 3; i++) num_doras += state.akas_in_hand[i] ? 1 : 0;
        num_doras += sup.num_doras_in_fuuro;

        Agari a = calc.agari(additional_yakus, num_doras);
        if (!a.valid) return false;
        if (a.is_yakuman) {
            float v = (float)a.point(is_oya).tsumo_total(is_oya);
            for (int i = 0; i < 4; i++) scores[i] = v;
            return true;
        }
        u8 fu = a.fu, han = a.han;
        for (int i = 0; i < 4; i++) scores[i] = 0.f;
        auto pts = [&](int h) {
            Agari x; x.valid = true; x.fu = fu; x.han = (u8)h;
            return (float)x.point(is_oya).tsumo_total(is_oya);
        };
        bool assume_riichi = sup.is_menzen && sup.prefer_riichi;
        if (assume_riichi && sup.n_dora_indicators == 1) {
            u8 n_indicators[5] = {0, 0, 0, 0, 0};
            u8 sum_indicators = 0;
            for (int tid = 0; tid < 34; tid++) {
                u8 count = state.tehai[tid];
                if (count == 0) continue;
                u8 ind_count = state.tiles_in_wall[tile_prev((u8)tid)];
                n_indicators[count] += ind_count;
                sum_indicators += ind_count;
            }
            float uradora_probs[5] = {0, 0, 0, 0, 0};
            u8 n_left = sum_left_tiles(state);
            uradora_probs[0] = (float)(u8)(n_left - sum_indicators) / (float)n_left;
            for (int i = 1; i < 5; i++) uradora_probs[i] = (float)n_indicators[i] / (float)n_left;
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 5; j++) {
                    float p = uradora_probs[j];
                    if (p == 0.f) continue;
                    scores[i] += pts(han + i + j) * p;
                }
        } else if (assume_riichi && sup.n_dora_indicators > 1) {
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 13; j++) {
                    float p = URADORA_PROB_TABLE[sup.n_dora_indicators - 1][j];
                    if (p == 0.f) continue;
                    scores[i] += pts(han + i + j) * p;
                }
        } else {
            for (int i = 0; i < 4; i++) scores[i] = pts(han + i);
        }
        return true;
    }
};

}  // namespace

static int cmp_f32_total(float a, float b) {
    // f32::total_cmp for the finite / non-negative values that occur here
    if (a < b) return -1;
    if (a > b) return 1;
    return 0;
}

// sp/candidate.rs:73-107
int sp_candidate_cmp(const SpCandidate& l, const SpCandidate& r, SpColumn by) {
    if (l.tile == r.tile) return 0;
    switch (by) {
        case SPCOL_EV: {
            int o = cmp_f32_total(l.exp_values.at(0), r.exp_values.at(0));
            return o != 0 ? o : sp_candidate_cmp(l, r, SPCOL_WIN_PROB);
        }
        case SPCOL_WIN_PROB: {
            int o = cmp_f32_total(l.win_probs.at(0), r.win_probs.at(0));
            return o != 0 ? o : sp_candidate_cmp(l, r, SPCOL_TENPAI_PROB);
        }
        case SPCOL_TENPAI_PROB: {
            int o = cmp_f32_total(l.tenpai_probs.at(0), r.tenpai_probs.at(0));
            return o != 0 ? o : sp_candidate_cmp(l, r, SPCOL_NOT_SHANTEN_DOWN);
        }
        case SPCOL_NOT_SHANTEN_DOWN:
            if (!l.shanten_down && r.shanten_down) return 1;
            if (l.shanten_down && !r.shanten_down) return -1;
            return sp_candidate_cmp(l, r, SPCOL_NUM_REQUIRED);
        case SPCOL_NUM_REQUIRED:
            if (l.num_required_tiles != r.num_required_tiles) return l.num_required_tiles < r.num_required_tiles ? -1 : 1;
            return sp_candidate_cmp(l, r, SPCOL_DISCARD_PRIORITY);
        case SPCOL_DISCARD_PRIORITY:
            return cmp_discard_priority(l.tile, r.tile);
    }
    return 0;
}

// sp/calc.rs:84-134
std::vector<SpCandidate> SpCalculator::calc(const SpInitState& init, bool can_discard, u8 tsumos_left,
                                            i8 cur_shanten) const {
    ORC_ENSURE(cur_shanten >= 0, "can't calculate an agari hand");
    ORC_ENSURE(tsumos_left >= 1, "need at least one more tsumo");
    ORC_ENSURE(tsumos_left <= MAX_TSUMOS_LEFT, "too many tsumos left");
    State st;
    memset(&st, 0, sizeof st);
    memcpy(st.tehai, init.tehai, 34);
    for (int i = 0; i < 3; i++) st.akas_in_hand[i] = init.akas_in_hand[i];
    for (int i = 0; i < 34; i++) st.tiles_in_wall[i] = 4 - init.tiles_seen[i];
    for (int i = 0; i < 3; i++) st.akas_in_wall[i] = !init.akas_seen[i];
    st.n_extra_tsumo = 0;
    int n_left = sum_left_tiles(st);
    CalcState cs(*this, st, tsumos_left, n_left);
    auto ret = cs.calc(can_discard, cur_shanten);
    long n = 0;
    for (int i = 0; i <= SHANTEN_THRES; i++) n += (long)cs.discard_cache[i].size() + (long)cs.draw_cache[i].size();
    g_sp_stats[0]++; g_sp_stats[1] += n; if (n > g_sp_stats[2]) g_sp_stats[2] = n;
    if (n > 1000) g_sp_stats[3]++;
    if (n > 10000) g_sp_stats[4]++;
    if (n > 100000) g_sp_stats[5]++;
    return ret;
}

// agent_helper.rs:509-593
bool single_player_tables(const PlayerState& ps, std::vector<SpCandidate>& out) {
    if (ps.tiles_left < 4) return false;
    i8 cur_shanten = ps.real_time_shanten();
    if (cur_shanten < 0) return false;

    bool can_discard = ps.last_cans.can_discard;
    u8 tsumos_left;
    bool calc_haitei;
    if (can_discard) {
        tsumos_left = ps.tiles_left / 4;
        calc_haitei = ps.tiles_left % 4 == 0;
    } else {
        u8 target = (u8)ps.rel(ps.last_cans.target_actor);
        u8 sub = 4 - target;
        u8 at_next = ps.tiles_left >= sub ? ps.tiles_left - sub : 0;
        tsumos_left = at_next / 4;
        calc_haitei = at_next % 4 == 0;
    }
    if (tsumos_left < 1) return false;

    u8 num_doras_in_fuuro;
    if (ps.is_menzen && ps.ankan_overview[0].empty()) {
        num_doras_in_fuuro = 0;
    } else {
        u8 in_tehai = 0;
        for (u8 ind : ps.dora_indicators) in_tehai += ps.tehai[tile_next(ind)];
        u8 num_akas = (u8)ps.akas_in_hand[0] + (u8)ps.akas_in_hand[1] + (u8)ps.akas_in_hand[2];
        num_doras_in_fuuro = ps.doras_owned[0] - in_tehai - num_akas;
    }
    bool prefer_riichi = ps.scores[0] >= 1000;
    bool calc_double_riichi = can_discard && ps.can_w_riichi;

    SpInitState init;
    memcpy(init.tehai, ps.tehai, 34);
    for (int i = 0; i < 3; i++) init.akas_in_hand[i] = ps.akas_in_hand[i];
    bool is_discard_after_riichi = can_discard && ps.riichi_accepted[0];
    if (is_discard_after_riichi) {
        u8 lt = ps.last_self_tsumo;
        init.tehai[deaka(lt)] -= 1;
        if (is_aka(lt)) init.akas_in_hand[lt - T_5MR] = false;
        can_discard = false;
    }
    memcpy(init.tiles_seen, ps.tiles_seen, 34);
    for (int i = 0; i < 3; i++) init.akas_seen[i] = ps.akas_seen[i];

    SpCalculator sp;
    sp.tehai_len_div3 = ps.tehai_len_div3;
    sp.is_menzen = ps.is_menzen;
    sp.chis = ps.chis.data(); sp.n_chis = (int)ps.chis.size();
    sp.pons = ps.pons.data(); sp.n_pons = (int)ps.pons.size();
    sp.minkans = ps.minkans.data(); sp.n_minkans = (int)ps.minkans.size();
    sp.ankans = ps.ankans.data(); sp.n_ankans = (int)ps.ankans.size();
    sp.bakaze = ps.bakaze; sp.jikaze = ps.jikaze;
    sp.num_doras_in_fuuro = num_doras_in_fuuro;
    sp.prefer_riichi = prefer_riichi;
    sp.dora_indicators = ps.dora_indicators.data(); sp.n_dora_indicators = (int)ps.dora_indicators.size();
    sp.calc_double_riichi = calc_double_riichi;
    sp.calc_haitei = calc_haitei;
    sp.sort_result = true;
    sp.maximize_win_prob = false;
    sp.calc_tegawari = false;
    sp.calc_shanten_down = false;

    try {
        out = sp.calc(init, can_discard, tsumos_left, cur_shanten);
    } catch (const OrcError&) {
        return false;
    }
    if (is_discard_after_riichi && !out.empty()) out[0].tile = ps.last_self_tsumo;
    return true;
}

}  // namespace orc
