// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// C ABI over the oracle for ctypes (tests/, smoke(), bench.py cpu_baseline / --impl reference).
#include "board.h"
#include "sp.h"

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

namespace orc { extern long g_sp_stats[8]; }
using namespace orc;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) { g_err = e.what(); return -1; }
}  // namespace

extern "C" {

// flat mirror of orc::Event for ctypes
struct orc_event {
    uint8_t type, actor, target, pai, tsumogiri;
    uint8_t consumed[4];
    uint8_t bakaze, kyoku, honba, kyotaku, oya;
    int32_t scores[4];
    uint8_t tehais[4][13];
    uint8_t has_deltas;
    int32_t deltas[4];
    uint8_t ura_markers[5];
    uint8_t n_ura;
};

static Event from_c(const orc_event& c) {
    Event e;
    e.type = c.type; e.actor = c.actor; e.target = c.target; e.pai = c.pai; e.tsumogiri = c.tsumogiri != 0;
    memcpy(e.consumed, c.consumed, 4);
    e.bakaze = c.bakaze; e.kyoku = c.kyoku; e.honba = c.honba; e.kyotaku = c.kyotaku; e.oya = c.oya;
    memcpy(e.scores, c.scores, sizeof e.scores);
    memcpy(e.tehais, c.tehais, sizeof e.tehais);
    e.has_deltas = c.has_deltas != 0;
    memcpy(e.deltas, c.deltas, sizeof e.deltas);
    memcpy(e.ura_markers, c.ura_markers, 5);
    e.n_ura = c.n_ura;
    return e;
}
static orc_event to_c(const Event& e) {
    orc_event c;
    memset(&c, 0, sizeof c);
    c.type = e.type; c.actor = e.actor; c.target = e.target; c.pai = e.pai; c.tsumogiri = e.tsumogiri;
    memcpy(c.consumed, e.consumed, 4);
    c.bakaze = e.bakaze; c.kyoku = e.kyoku; c.honba = e.honba; c.kyotaku = e.kyotaku; c.oya = e.oya;
    memcpy(c.scores, e.scores, sizeof c.scores);
    memcpy(c.tehais, e.tehais, sizeof c.tehais);
    c.has_deltas = e.has_deltas;
    memcpy(c.deltas, e.deltas, sizeof c.deltas);
    memcpy(c.ura_markers, e.ura_markers, 5);
    c.n_ura = (uint8_t)e.n_ura;
    return c;
}

const char* orc_last_error() { return g_err.c_str(); }

int orc_init(const char* data_dir) {
    try { tables_init(data_dir); return 0; } catch (const std::exception& e) { return fail(e); }
}

// ---------------- algo ----------------
int orc_shanten(const uint8_t* tiles, const uint8_t* len_div3, int8_t* out, int n, int kind) {
    try {
        for (int i = 0; i < n; i++) {
            const u8* t = tiles + (size_t)i * 34;
            out[i] = kind == 0 ? shanten_all(t, len_div3[i]) : kind == 1 ? shanten_normal(t, len_div3[i])
                     : kind == 2 ? shanten_chitoi(t) : shanten_kokushi(t);
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// one agari query; layout shared with include/mjx.h mjx_agari_in / mjx_agari_out
struct orc_agari_in {
    uint8_t tehai[34];
    uint8_t chis[4], pons[4], minkans[4], ankans[4];
    uint8_t n_chis, n_pons, n_minkans, n_ankans;
    uint8_t bakaze, jikaze, winning_tile, is_ron;
    uint8_t additional_hans, doras;  // for agari(); ignored by search_yakus
    uint8_t is_oya, pad;
};
struct orc_agari_out {
    uint8_t kind;  // 0 none, 1 normal, 2 yakuman
    uint8_t fu, han, yakuman;
    int32_t ron, tsumo_ko, tsumo_oya;  // Point for is_oya (0 if none / panic combo -> -1)
};

// mode 0: search_yakus, 1: agari(additional_hans, doras), 2: has_yaku (kind=1 if true)
int orc_agari(const orc_agari_in* in, orc_agari_out* out, int n, int mode) {
    try {
        for (int i = 0; i < n; i++) {
            const orc_agari_in& q = in[i];
            AgariCalc c;
            c.tehai = q.tehai;
            c.chis = q.chis; c.n_chis = q.n_chis;
            c.pons = q.pons; c.n_pons = q.n_pons;
            c.minkans = q.minkans; c.n_minkans = q.n_minkans;
            c.ankans = q.ankans; c.n_ankans = q.n_ankans;
            c.is_menzen = q.n_chis == 0 && q.n_pons == 0 && q.n_minkans == 0;
            c.bakaze = q.bakaze; c.jikaze = q.jikaze; c.winning_tile = q.winning_tile; c.is_ron = q.is_ron;
            orc_agari_out& o = out[i];
            memset(&o, 0, sizeof o);
            if (mode == 2) { o.kind = c.has_yaku() ? 1 : 0; continue; }
            Agari a = mode == 0 ? c.search_yakus() : c.agari(q.additional_hans, q.doras);
            if (!a.valid) continue;
            o.kind = a.is_yakuman ? 2 : 1;
            o.fu = a.fu; o.han = a.han; o.yakuman = a.yakuman;
            try {
                Point p = a.point(q.is_oya);
                o.ron = p.ron; o.tsumo_ko = p.tsumo_ko; o.tsumo_oya = p.tsumo_oya;
            } catch (const OrcError&) {
                o.ron = o.tsumo_ko = o.tsumo_oya = -1;
            }
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

int orc_point(int is_oya, int fu, int han, int32_t* out3) {
    try {
        Point p = point_calc(is_oya, (u8)fu, (u8)han);
        out3[0] = p.ron; out3[1] = p.tsumo_ko; out3[2] = p.tsumo_oya;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

int orc_check_ankan_after_riichi(const uint8_t* tehai, int len_div3, int tile, int strict) {
    try { return check_ankan_after_riichi(tehai, (u8)len_div3, (u8)tile, strict) ? 1 : 0; }
    catch (const std::exception& e) { return fail(e); }
}

void orc_rankings(const int32_t* scores, uint8_t* player_by_rank, uint8_t* rank_by_player) {
    rankings(scores, player_by_rank, rank_by_player);
}

uint32_t orc_agari_key(const uint8_t* tiles, uint8_t* tile14) { return get_tile14_and_key(tiles, tile14); }
int orc_agari_lookup(uint32_t key, uint32_t* divs4) { return agari_table_lookup(key, divs4); }

// ---------------- wall ----------------
void orc_make_wall(uint64_t nonce, uint64_t key, int kyoku, int honba, int shuffle_kind, uint8_t* seq136) {
    make_wall(nonce, key, (u8)kyoku, (u8)honba, shuffle_kind, seq136);
}
void orc_sha3_256(const uint8_t* data, int len, uint8_t* out32) { sha3_256(data, (size_t)len, out32); }
void orc_chacha12(const uint8_t* seed32, uint32_t* out, int n) {
    ChaCha12 r(seed32);
    for (int i = 0; i < n; i++) out[i] = r.next_u32();
}

// ---------------- PlayerState ----------------
void* orc_ps_new(int player_id) { return new PlayerState((u8)player_id); }
void orc_ps_free(void* p) { delete static_cast<PlayerState*>(p); }
void* orc_ps_clone(void* p) { return new PlayerState(*static_cast<PlayerState*>(p)); }

static uint32_t pack_cans(const ActionCandidate& c) {
    uint32_t v = 0;
    v |= (uint32_t)c.can_discard << 0; v |= (uint32_t)c.can_chi_low << 1; v |= (uint32_t)c.can_chi_mid << 2;
    v |= (uint32_t)c.can_chi_high << 3; v |= (uint32_t)c.can_pon << 4; v |= (uint32_t)c.can_daiminkan << 5;
    v |= (uint32_t)c.can_kakan << 6; v |= (uint32_t)c.can_ankan << 7; v |= (uint32_t)c.can_riichi << 8;
    v |= (uint32_t)c.can_tsumo_agari << 9; v |= (uint32_t)c.can_ron_agari << 10; v |= (uint32_t)c.can_ryukyoku << 11;
    v |= (uint32_t)c.target_actor << 16;
    return v;
}

// returns packed cans (>=0) or -1
int64_t orc_ps_update(void* p, const orc_event* ev) {
    try { return pack_cans(static_cast<PlayerState*>(p)->update(from_c(*ev))); }
    catch (const std::exception& e) { return fail(e); }
}
int orc_ps_validate_reaction(void* p, const orc_event* ev) {
    try { static_cast<PlayerState*>(p)->validate_reaction(from_c(*ev)); return 0; }
    catch (const std::exception& e) { return fail(e); }
}

// field snapshot used by tests and by GPU-vs-oracle state diffs
struct orc_ps_view {
    uint8_t tehai[34], waits[34], dora_factor[34], tiles_seen[34], keep_shanten_discards[34],
        next_shanten_discards[34], forbidden_tiles[34], discarded_tiles[34];
    uint8_t akas_seen[3], akas_in_hand[3];
    uint8_t bakaze, jikaze, kyoku, honba, kyotaku, rank, oya, is_all_last;
    int32_t scores[4];
    uint8_t n_dora_indicators, dora_indicators[5];
    uint8_t riichi_declared[4], riichi_accepted[4];
    uint8_t at_turn, tiles_left;
    int8_t shanten, real_time_shanten;
    uint8_t has_last_self_tsumo, last_self_tsumo, has_last_kawa_tile, last_kawa_tile;
    uint32_t cans;
    uint8_t n_ankan_candidates, ankan_candidates[3], n_kakan_candidates, kakan_candidates[3];
    uint8_t chankan_chance, can_w_riichi, is_w_riichi, at_rinshan, at_ippatsu, at_furiten,
        to_mark_same_cycle_furiten, kans_on_board, is_menzen;
    uint8_t n_chis, chis[4], n_pons, pons[4], n_minkans, minkans[4], n_ankans, ankans[4];
    uint8_t doras_owned[4], doras_seen, tehai_len_div3, has_next_shanten_discard;
    uint8_t kawa_len[4];
};

void orc_ps_view_get(void* p, orc_ps_view* v) {
    const PlayerState& s = *static_cast<PlayerState*>(p);
    memset(v, 0, sizeof *v);
    for (int i = 0; i < 34; i++) {
        v->tehai[i] = s.tehai[i]; v->waits[i] = s.waits[i]; v->dora_factor[i] = s.dora_factor[i];
        v->tiles_seen[i] = s.tiles_seen[i]; v->keep_shanten_discards[i] = s.keep_shanten_discards[i];
        v->next_shanten_discards[i] = s.next_shanten_discards[i]; v->forbidden_tiles[i] = s.forbidden_tiles[i];
        v->discarded_tiles[i] = s.discarded_tiles[i];
    }
    for (int i = 0; i < 3; i++) { v->akas_seen[i] = s.akas_seen[i]; v->akas_in_hand[i] = s.akas_in_hand[i]; }
    v->bakaze = s.bakaze; v->jikaze = s.jikaze; v->kyoku = s.kyoku; v->honba = s.honba; v->kyotaku = s.kyotaku;
    v->rank = s.rank; v->oya = s.oya; v->is_all_last = s.is_all_last;
    for (int i = 0; i < 4; i++) v->scores[i] = s.scores[i];
    v->n_dora_indicators = (u8)s.dora_indicators.size();
    for (size_t i = 0; i < s.dora_indicators.size() && i < 5; i++) v->dora_indicators[i] = s.dora_indicators[i];
    for (int i = 0; i < 4; i++) { v->riichi_declared[i] = s.riichi_declared[i]; v->riichi_accepted[i] = s.riichi_accepted[i]; }
    v->at_turn = s.at_turn; v->tiles_left = s.tiles_left; v->shanten = s.shanten;
    v->real_time_shanten = s.real_time_shanten();
    v->has_last_self_tsumo = s.has_last_self_tsumo; v->last_self_tsumo = s.last_self_tsumo;
    v->has_last_kawa_tile = s.has_last_kawa_tile; v->last_kawa_tile = s.last_kawa_tile;
    v->cans = pack_cans(s.last_cans);
    v->n_ankan_candidates = (u8)s.ankan_candidates.size();
    for (size_t i = 0; i < s.ankan_candidates.size() && i < 3; i++) v->ankan_candidates[i] = s.ankan_candidates[i];
    v->n_kakan_candidates = (u8)s.kakan_candidates.size();
    for (size_t i = 0; i < s.kakan_candidates.size() && i < 3; i++) v->kakan_candidates[i] = s.kakan_candidates[i];
    v->chankan_chance = s.chankan_chance; v->can_w_riichi = s.can_w_riichi; v->is_w_riichi = s.is_w_riichi;
    v->at_rinshan = s.at_rinshan; v->at_ippatsu = s.at_ippatsu; v->at_furiten = s.at_furiten;
    v->to_mark_same_cycle_furiten = s.to_mark_same_cycle_furiten; v->kans_on_board = s.kans_on_board;
    v->is_menzen = s.is_menzen;
    auto cp = [](const std::vector<u8>& src, uint8_t* n, uint8_t* dst) {
        *n = (u8)src.size();
        for (size_t i = 0; i < src.size() && i < 4; i++) dst[i] = src[i];
    };
    cp(s.chis, &v->n_chis, v->chis); cp(s.pons, &v->n_pons, v->pons);
    cp(s.minkans, &v->n_minkans, v->minkans); cp(s.ankans, &v->n_ankans, v->ankans);
    for (int i = 0; i < 4; i++) { v->doras_owned[i] = s.doras_owned[i]; v->kawa_len[i] = (u8)s.kawa[i].size(); }
    v->doras_seen = s.doras_seen; v->tehai_len_div3 = s.tehai_len_div3;
    v->has_next_shanten_discard = s.has_next_shanten_discard;
}

// direct field pokes for the unit tests that construct states by hand (state/test.rs:70-220)
void orc_ps_set_tehai(void* p, const uint8_t* tehai34, int len_div3) {
    PlayerState& s = *static_cast<PlayerState*>(p);
    memcpy(s.tehai, tehai34, 34);
    s.tehai_len_div3 = (u8)len_div3;
}
int orc_ps_update_waits_and_furiten(void* p) {
    try {
        PlayerState& s = *static_cast<PlayerState*>(p);
        s.update_shanten();
        s.update_waits_and_furiten();
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
uint32_t orc_ps_set_can_chi_from_tile(void* p, int tile) {
    PlayerState& s = *static_cast<PlayerState*>(p);
    s.set_can_chi_from_tile((u8)tile);
    return pack_cans(s.last_cans);
}
int orc_ps_get_rank(int player_id, const int32_t* scores_rel) {
    PlayerState s((u8)player_id);
    return s.get_rank(scores_rel);
}

int orc_ps_agari_points(void* p, int is_ron, const uint8_t* ura, int n_ura, int32_t* out3) {
    try {
        Point pt = static_cast<PlayerState*>(p)->agari_points(is_ron != 0, ura, n_ura);
        out3[0] = pt.ron; out3[1] = pt.tsumo_ko; out3[2] = pt.tsumo_oya;
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
int orc_ps_rule_based_agari(void* p) {
    try { return static_cast<PlayerState*>(p)->rule_based_agari() ? 1 : 0; }
    catch (const std::exception& e) { return fail(e); }
}
int orc_ps_rule_based_agari_slow(void* p, int is_ron, int target_rel) {
    try { return static_cast<PlayerState*>(p)->rule_based_agari_slow(is_ron != 0, target_rel) ? 1 : 0; }
    catch (const std::exception& e) { return fail(e); }
}
int orc_ps_discard_candidates(void* p, int kind, uint8_t* out37) {
    try {
        bool b[37];
        PlayerState& s = *static_cast<PlayerState*>(p);
        if (kind == 0) s.discard_candidates_aka(b);
        else s.discard_candidates_with_unconditional_tenpai_aka(b);
        for (int i = 0; i < 37; i++) out37[i] = b[i];
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
int orc_obs_rows(int version) { try { return obs_rows(version); } catch (const std::exception& e) { return fail(e); } }
int orc_ps_encode_obs(void* p, int version, int at_kan_select, float* obs, uint8_t* mask46, int sp_mode) {
    try { static_cast<PlayerState*>(p)->encode_obs(version, at_kan_select != 0, obs, mask46, sp_mode); return 0; }
    catch (const std::exception& e) { return fail(e); }
}
int orc_ps_legal_mask(void* p, int at_kan_select, uint8_t* mask46) {
    try { legal_mask(*static_cast<PlayerState*>(p), at_kan_select != 0, mask46); return 0; }
    catch (const std::exception& e) { return fail(e); }
}

// ---------------- SP (algo/sp) ----------------
struct orc_sp_in {
    uint8_t tehai[34], akas_in_hand[3], tiles_seen[34], akas_seen[3];
    uint8_t tehai_len_div3, is_menzen, bakaze, jikaze, num_doras_in_fuuro;
    uint8_t n_dora_indicators, dora_indicators[5];
    uint8_t calc_double_riichi, calc_haitei, prefer_riichi, sort_result, maximize_win_prob, calc_tegawari,
        calc_shanten_down;
    uint8_t chis[4], pons[4], minkans[4], ankans[4], n_chis, n_pons, n_minkans, n_ankans;
    uint8_t can_discard, tsumos_left;
    int8_t cur_shanten;
};
struct orc_sp_cand {
    uint8_t tile, shanten_down, num_required_tiles, n_required, n_turns;
    uint8_t required_tile[34], required_count[34];
    float tenpai_probs[17], win_probs[17], exp_values[17];
};
int orc_sp_calc(const orc_sp_in* q, orc_sp_cand* out, int max_out) {
    try {
        SpCalculator sp;
        sp.tehai_len_div3 = q->tehai_len_div3; sp.is_menzen = q->is_menzen; sp.bakaze = q->bakaze; sp.jikaze = q->jikaze;
        sp.num_doras_in_fuuro = q->num_doras_in_fuuro;
        sp.dora_indicators = q->dora_indicators; sp.n_dora_indicators = q->n_dora_indicators;
        sp.calc_double_riichi = q->calc_double_riichi; sp.calc_haitei = q->calc_haitei; sp.prefer_riichi = q->prefer_riichi;
        sp.sort_result = q->sort_result; sp.maximize_win_prob = q->maximize_win_prob; sp.calc_tegawari = q->calc_tegawari;
        sp.calc_shanten_down = q->calc_shanten_down;
        sp.chis = q->chis; sp.n_chis = q->n_chis; sp.pons = q->pons; sp.n_pons = q->n_pons;
        sp.minkans = q->minkans; sp.n_minkans = q->n_minkans; sp.ankans = q->ankans; sp.n_ankans = q->n_ankans;
        SpInitState init;
        memcpy(init.tehai, q->tehai, 34);
        memcpy(init.tiles_seen, q->tiles_seen, 34);
        for (int i = 0; i < 3; i++) { init.akas_in_hand[i] = q->akas_in_hand[i]; init.akas_seen[i] = q->akas_seen[i]; }
        auto cands = sp.calc(init, q->can_discard, q->tsumos_left, q->cur_shanten);
        int n = 0;
        for (auto& c : cands) {
            if (n >= max_out) break;
            orc_sp_cand& o = out[n++];
            memset(&o, 0, sizeof o);
            o.tile = c.tile; o.shanten_down = c.shanten_down; o.num_required_tiles = c.num_required_tiles;
            o.n_required = (u8)c.required_tiles.size();
            for (size_t i = 0; i < c.required_tiles.size(); i++) {
                o.required_tile[i] = c.required_tiles[i].tile; o.required_count[i] = c.required_tiles[i].count;
            }
            o.n_turns = (u8)c.tenpai_probs.size();
            for (size_t i = 0; i < c.tenpai_probs.size(); i++) {
                o.tenpai_probs[i] = c.tenpai_probs[i]; o.win_probs[i] = c.win_probs[i]; o.exp_values[i] = c.exp_values[i];
            }
        }
        return n;
    } catch (const std::exception& e) { return fail(e); }
}

// ---------------- manual game driving (golden-log replay) ----------------
struct OrcGame {
    Game g;
    std::vector<u8> walls;  // optional injected walls, keyed sequentially
};

void* orc_game_new(uint64_t nonce, uint64_t key, int shuffle_kind, int table) {
    OrcGame* og = new OrcGame();
    og->g.seed_nonce = nonce; og->g.seed_key = key; og->g.shuffle_kind = shuffle_kind; og->g.table = table;
    return og;
}
void orc_game_free(void* p) { delete static_cast<OrcGame*>(p); }
// returns 1 if ended, 0 in game, -1 error
int orc_game_poll(void* p) {
    try { Game& g = static_cast<OrcGame*>(p)->g; g.poll(); return g.ended ? 1 : 0; }
    catch (const std::exception& e) { return fail(e); }
}
void* orc_game_state(void* p, int seat) {
    Game& g = static_cast<OrcGame*>(p)->g;
    return g.board ? &g.board->player_states[seat] : nullptr;
}
void orc_game_set_reaction(void* p, int seat, const orc_event* ev) {
    static_cast<OrcGame*>(p)->g.last_reactions[seat] = from_c(*ev);
}
// decode an action id for `seat` the way MortalBatchAgent does and stage it as the reaction
int orc_game_set_action(void* p, int seat, int action, int kan_action) {
    try {
        Game& g = static_cast<OrcGame*>(p)->g;
        g.last_reactions[seat] = decode_action(g.board->player_states[seat], (u8)seat, action, kan_action);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
int orc_oracle_obs_rows(int version) { try { return oracle_obs_rows(version); } catch (const std::exception& e) { return fail(e); } }
// board.rs:680-782 encode_oracle_obs of the running kyoku from `perspective`
int orc_game_encode_oracle_obs(void* p, int perspective, int version, float* out) {
    try {
        Game& g = static_cast<OrcGame*>(p)->g;
        if (!g.board) throw OrcError("no running kyoku");
        g.board->encode_oracle_obs((u8)perspective, version, out);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
void orc_game_advance_step(void* p) { static_cast<OrcGame*>(p)->g.step_idx++; }
// finalises scores once the game has ended (game.rs:180-184)
int orc_game_finish(void* p, int32_t* scores4) {
    try {
        Game& g = static_cast<OrcGame*>(p)->g;
        AgentConfig cfg[4]; PolicyFn pol[4];
        if (!g.ended) throw OrcError("game not ended");
        g.commit(cfg, pol, nullptr);
        for (int i = 0; i < 4; i++) scores4[i] = g.scores[i];
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
void orc_game_info(void* p, int32_t* out /* kyoku, honba, kyotaku, scores[4], ended, n_kyoku_logs */) {
    Game& g = static_cast<OrcGame*>(p)->g;
    out[0] = g.kyoku; out[1] = g.honba; out[2] = g.kyotaku;
    for (int i = 0; i < 4; i++) out[3 + i] = g.scores[i];
    out[7] = g.ended; out[8] = (int32_t)g.game_log.size();
}
// copies the event log: finished kyokus followed by the running kyoku's log
int orc_game_log(void* p, orc_event* out, int max_out) {
    Game& g = static_cast<OrcGame*>(p)->g;
    int n = 0;
    for (auto& k : g.game_log)
        for (auto& e : k) { if (n < max_out) out[n] = to_c(e); n++; }
    if (g.board && g.kyoku_started)
        for (auto& e : g.board->log) { if (n < max_out) out[n] = to_c(e); n++; }
    return n;
}

// ---------------- batch self-play with built-in policies ----------------
// Per-decision trace record: [table, step_idx, seat, action, kan_action, mask_lo, mask_hi]
struct orc_run_cfg {
    int32_t n_tables;
    int32_t shuffle_kind;
    int32_t policy_kind;        // test_policy kind for all seats
    int32_t enable_quick_eval;
    int32_t enable_agari_guard;
    int32_t encode_obs;         // 0 none; else obs version to encode for every policy row (timed work)
    int32_t sp_mode;
    int32_t n_threads;
    int64_t max_steps_per_table;  // stop early after this many table-steps (0 = run to the end)
    int64_t encode_from_step;     // bench fast-forward: steps before this one are played without encoding and not counted
};
struct orc_run_out {
    int64_t table_steps;   // sum over tables of decision cycles (game.rs:304 `actions`)
    int64_t obs_rows;      // rows handed to the policy (incl. kan-select rows)
    double seconds;
};

// One table's run, kept alive across the two phases of a batch (fast-forward, then the timed / sampled part).
struct TableRun {
    Game g;
    int64_t n_steps = 0, rows = 0;
    bool done = false;
};

// Optional observation sampling (parity at benchmark scale): `samples` = int64 [m, 4] rows (table, step_idx, seat, kan_select)
// sorted lexicographically; the observation / mask of every decision found there is written to obs_out[idx] / masks_out[idx].
struct SampleSink {
    const int64_t* samples = nullptr;
    int64_t m = 0;
    int version = 4;
    float* obs_out = nullptr;
    float* inv_out = nullptr;  // optional: the invisible (oracle) observation of the same decisions
    uint8_t* masks_out = nullptr;
    uint8_t* found = nullptr;
    int64_t find(const Scene& sc) const {
        const int64_t key[4] = {sc.table, (int64_t)sc.step_idx, sc.seat, sc.is_kan_select ? 1 : 0};
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            const int64_t* r = samples + mid * 4;
            bool less = false;
            for (int q = 0; q < 4; q++) { if (r[q] != key[q]) { less = r[q] < key[q]; break; } }
            if (less) lo = mid + 1; else hi = mid;
        }
        if (lo >= m) return -1;
        const int64_t* r = samples + lo * 4;
        return (r[0] == key[0] && r[1] == key[1] && r[2] == key[2] && r[3] == key[3]) ? lo : -1;
    }
};

// Advance table t until it ends or has played `until` table-steps (0 = to the end). Rows at step_idx >= encode_from are
// encoded when cfg.encode_obs is set (the timed work of the CPU arm).
static void run_table(const orc_run_cfg& cfg, TableRun& tr, int t, int64_t until, int64_t encode_from, std::vector<float>& obs,
                      int64_t* trace, int64_t trace_cap, std::atomic<int64_t>* trace_len, const SampleSink* sink) {
    Game& g = tr.g;
    AgentConfig ac;
    ac.enable_quick_eval = cfg.enable_quick_eval != 0;
    ac.enable_rule_based_agari_guard = cfg.enable_agari_guard != 0;
    AgentConfig cfgs[4] = {ac, ac, ac, ac};
    PolicyFn pol = [&](const Scene& sc, const u8* mask, float*) {
        tr.rows++;
        if (cfg.encode_obs && (int64_t)sc.step_idx >= encode_from) {
            u8 m2[46];
            sc.state->encode_obs(cfg.encode_obs, sc.is_kan_select, obs.data(), m2, cfg.sp_mode);
        }
        if (sink && sink->m) {
            const int64_t at = sink->find(sc);
            if (at >= 0) {
                const size_t stride = (size_t)obs_rows(sink->version) * 34;
                sc.state->encode_obs(sink->version, sc.is_kan_select, sink->obs_out + (size_t)at * stride,
                                     sink->masks_out + (size_t)at * 46, cfg.sp_mode);
                if (sink->inv_out && sc.board)
                    sc.board->encode_oracle_obs(sc.seat, sink->version, sink->inv_out + (size_t)at * oracle_obs_rows(sink->version) * 34);
                sink->found[at] = 1;
            }
        }
        int a = test_policy(cfg.policy_kind, sc, g.seed_nonce, g.seed_key, mask);
        if (trace) {
            int64_t at = trace_len->fetch_add(1);
            if (at < trace_cap) {
                u64 bits = 0;
                for (int i = 0; i < 46; i++) if (mask[i]) bits |= 1ull << i;
                int64_t* r = trace + at * 6;
                r[0] = sc.table; r[1] = (int64_t)sc.step_idx; r[2] = sc.seat; r[3] = a;
                r[4] = sc.is_kan_select; r[5] = (int64_t)bits;
            }
        }
        return a;
    };
    PolicyFn pols[4] = {pol, pol, pol, pol};
    while (!tr.done && (until <= 0 || tr.n_steps < until)) {
        g.poll();
        if (g.commit(cfgs, pols, nullptr)) { tr.done = true; break; }
        tr.n_steps++;
    }
    (void)t;
}

// Two phases, each with dynamic table hand-out over n_threads workers: (1) untimed fast-forward of every table to
// `encode_from_step` table-steps (skipped when 0), (2) the timed part up to `max_steps_per_table` (0 = to the end).
// `seconds` clocks phase 2 only; `table_steps` counts phase-2 steps only.
static int run_batch_impl(const orc_run_cfg* cfg, const uint64_t* nonces, const uint64_t* keys, const int32_t* table_ids,
                          int32_t* scores, uint8_t* ranks, int32_t* steps, int64_t* trace, int64_t trace_cap,
                          int64_t* trace_len_out, orc_run_out* out, const SampleSink* sink) {
    try {
        const int n = cfg->n_tables;
        const int nt = std::max(1, cfg->n_threads);
        std::atomic<int64_t> tlen(0);
        std::vector<std::unique_ptr<TableRun>> runs(n);
        for (int t = 0; t < n; t++) {
            runs[t].reset(new TableRun());
            Game& g = runs[t]->g;
            g.seed_nonce = nonces[t]; g.seed_key = keys[t]; g.shuffle_kind = cfg->shuffle_kind;
            g.table = table_ids ? table_ids[t] : t;
        }
        std::vector<std::string> errs(nt);
        auto phase = [&](int64_t until, int64_t encode_from) {
            std::atomic<int> next(0);
            std::vector<std::thread> th;
            for (int k = 0; k < nt; k++)
                th.emplace_back([&, k]() {
                    try {
                        std::vector<float> obs;
                        if (cfg->encode_obs) obs.resize((size_t)obs_rows(cfg->encode_obs) * 34);
                        for (;;) {
                            const int t = next.fetch_add(1);
                            if (t >= n) break;
                            run_table(*cfg, *runs[t], t, until, encode_from, obs, trace, trace_cap, &tlen, sink);
                        }
                    } catch (const std::exception& e) { errs[k] = e.what(); }
                });
            for (auto& t : th) t.join();
            for (auto& e : errs) if (!e.empty()) throw OrcError(e);
        };
        const int64_t ff = cfg->encode_from_step;
        if (ff > 0) phase(ff, ff);
        int64_t before = 0;
        for (int t = 0; t < n; t++) before += runs[t]->n_steps;
        auto t0 = std::chrono::steady_clock::now();
        phase(cfg->max_steps_per_table, ff);
        auto t1 = std::chrono::steady_clock::now();
        int64_t total = 0, rows = 0;
        for (int t = 0; t < n; t++) {
            const Game& g = runs[t]->g;
            for (int i = 0; i < 4; i++) scores[t * 4 + i] = g.scores[i];
            rankings(g.scores, nullptr, ranks + t * 4);
            steps[t] = (int32_t)runs[t]->n_steps;
            total += runs[t]->n_steps;
            rows += runs[t]->rows;
        }
        if (trace_len_out) *trace_len_out = tlen.load();
        if (out) {
            out->table_steps = total - before;
            out->obs_rows = rows;
            out->seconds = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- persistent batch (bench.py --impl reference): the tables live across calls, so one timed "step" is one pass
// of the whole batch, exactly like BatchGame::run's loop body (game.rs:286-304).
struct OrcBatch {
    orc_run_cfg cfg;
    std::vector<std::unique_ptr<TableRun>> runs;
};
void* orc_batch_new(const orc_run_cfg* cfg, const uint64_t* nonces, const uint64_t* keys) {
    OrcBatch* b = new OrcBatch();
    b->cfg = *cfg;
    b->runs.resize(cfg->n_tables);
    for (int t = 0; t < cfg->n_tables; t++) {
        b->runs[t].reset(new TableRun());
        Game& g = b->runs[t]->g;
        g.seed_nonce = nonces[t]; g.seed_key = keys[t]; g.shuffle_kind = cfg->shuffle_kind; g.table = t;
    }
    return b;
}
void orc_batch_free(void* p) { delete static_cast<OrcBatch*>(p); }
// advance every live table to `until` table-steps (dynamic hand-out over cfg.n_threads workers); rows at step_idx >=
// encode_from are encoded when cfg.encode_obs is set. out: table-steps advanced by this call, rows, seconds of this call.
int orc_batch_run(void* p, int64_t until, int64_t encode_from, orc_run_out* out) {
    try {
        OrcBatch& b = *static_cast<OrcBatch*>(p);
        const int n = b.cfg.n_tables, nt = std::max(1, b.cfg.n_threads);
        int64_t before = 0, rows_before = 0;
        for (int t = 0; t < n; t++) { before += b.runs[t]->n_steps; rows_before += b.runs[t]->rows; }
        std::vector<std::string> errs(nt);
        std::atomic<int> next(0);
        std::atomic<int64_t> tlen(0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int k = 0; k < nt; k++)
            th.emplace_back([&, k]() {
                try {
                    std::vector<float> obs;
                    if (b.cfg.encode_obs) obs.resize((size_t)obs_rows(b.cfg.encode_obs) * 34);
                    for (;;) {
                        const int t = next.fetch_add(1);
                        if (t >= n) break;
                        run_table(b.cfg, *b.runs[t], t, until, encode_from, obs, nullptr, 0, &tlen, nullptr);
                    }
                } catch (const std::exception& e) { errs[k] = e.what(); }
            });
        for (auto& t : th) t.join();
        auto t1 = std::chrono::steady_clock::now();
        for (auto& e : errs) if (!e.empty()) throw OrcError(e);
        int64_t after = 0, rows_after = 0;
        for (int t = 0; t < n; t++) { after += b.runs[t]->n_steps; rows_after += b.runs[t]->rows; }
        if (out) {
            out->table_steps = after - before;
            out->obs_rows = rows_after - rows_before;
            out->seconds = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

int orc_run_batch(const orc_run_cfg* cfg, const uint64_t* nonces, const uint64_t* keys, const int32_t* table_ids,
                  int32_t* scores, uint8_t* ranks, int32_t* steps, int64_t* trace, int64_t trace_cap,
                  int64_t* trace_len_out, orc_run_out* out) {
    return run_batch_impl(cfg, nonces, keys, table_ids, scores, ranks, steps, trace, trace_cap, trace_len_out, out, nullptr);
}

// The same run, additionally encoding the decisions listed in `samples` (see SampleSink) with obs `version`.
int orc_run_sample_obs(const orc_run_cfg* cfg, const uint64_t* nonces, const uint64_t* keys, int32_t* scores, uint8_t* ranks,
                       int32_t* steps, const int64_t* samples, int64_t m, int version, float* obs_out, uint8_t* masks_out,
                       uint8_t* found, float* inv_out) {
    SampleSink sink;
    sink.samples = samples; sink.m = m; sink.version = version; sink.obs_out = obs_out; sink.masks_out = masks_out; sink.found = found;
    sink.inv_out = inv_out;
    return run_batch_impl(cfg, nonces, keys, nullptr, scores, ranks, steps, nullptr, 0, nullptr, nullptr, &sink);
}

void orc_sp_stats(long* out) { for (int i = 0; i < 8; i++) { out[i] = orc::g_sp_stats[i]; orc::g_sp_stats[i] = 0; } }

// Action-replay parity (SURVEY.md §8d protocol ii): drive the oracle with decisions recorded elsewhere.
// replay rows: [table, step_idx, seat, kan_select, action], sorted lexicographically by the first four.
// Every replayed action must be legal in the oracle's own mask, otherwise the call fails. `mask_bits` (optional, aligned
// with the replay rows) = the legal mask the recorder saw, compared with the oracle's bit for bit. `max_steps` > 0 stops
// every table after that many table-steps (the recording was cut at the same point); scores are then the running scores.
static const uint8_t* g_replay_qe_flags = nullptr;  // optional [n_tables, 4] per-seat enable_quick_eval (set by orc_run_replay3)
int orc_run_replay2(int n_tables, const uint64_t* nonces, const uint64_t* keys, int shuffle_kind, int enable_quick_eval,
                    const int64_t* replay, int64_t n_replay, const int64_t* mask_bits, int64_t max_steps, int n_threads,
                    int32_t* scores, uint8_t* ranks, int32_t* steps) {
    const uint8_t* qe_flags = g_replay_qe_flags;
    try {
        std::atomic<int64_t> used(0);
        std::atomic<int> next(0);
        const int nt = std::max(1, n_threads);
        std::vector<std::string> errs(nt);
        auto work = [&](int k) {
            try {
                for (;;) {
                    const int t = next.fetch_add(1);
                    if (t >= n_tables) break;
                    Game g;
                    g.seed_nonce = nonces[t]; g.seed_key = keys[t]; g.shuffle_kind = shuffle_kind; g.table = t;
                    AgentConfig ac;
                    ac.enable_quick_eval = enable_quick_eval != 0;
                    AgentConfig cfgs[4] = {ac, ac, ac, ac};
                    if (qe_flags) for (int sx = 0; sx < 4; sx++) cfgs[sx].enable_quick_eval = qe_flags[t * 4 + sx] != 0;  // mortal.rs:54-74
                    PolicyFn pol = [&](const Scene& sc, const u8* mask, float*) {
                        int64_t key[4] = {sc.table, (int64_t)sc.step_idx, sc.seat, sc.is_kan_select ? 1 : 0};
                        int64_t lo = 0, hi = n_replay;
                        while (lo < hi) {
                            int64_t mid = (lo + hi) / 2;
                            const int64_t* r = replay + mid * 5;
                            bool less = false;
                            for (int q = 0; q < 4; q++) { if (r[q] != key[q]) { less = r[q] < key[q]; break; } }
                            if (less) lo = mid + 1; else hi = mid;
                        }
                        const int64_t* r = replay + lo * 5;
                        if (lo >= n_replay || r[0] != key[0] || r[1] != key[1] || r[2] != key[2] || r[3] != key[3])
                            throw OrcError("replay: no recorded decision for table " + std::to_string(sc.table) + " step " +
                                           std::to_string(sc.step_idx) + " seat " + std::to_string(sc.seat));
                        int a = (int)r[4];
                        if (a < 0 || a >= 46 || !mask[a]) throw OrcError("replay: recorded action is illegal in the oracle");
                        if (mask_bits) {
                            u64 bits = 0;
                            for (int i = 0; i < 46; i++) if (mask[i]) bits |= 1ull << i;
                            if ((int64_t)bits != mask_bits[lo])
                                throw OrcError("replay: recorded legal mask differs from the oracle's at table " + std::to_string(sc.table) +
                                               " step " + std::to_string(sc.step_idx) + " seat " + std::to_string(sc.seat));
                        }
                        used.fetch_add(1);
                        return a;
                    };
                    PolicyFn pols[4] = {pol, pol, pol, pol};
                    int64_t n_steps = 0;
                    for (;;) {
                        g.poll();
                        if (g.commit(cfgs, pols, nullptr)) break;
                        n_steps++;
                        if (max_steps > 0 && n_steps >= max_steps) break;
                    }
                    for (int i = 0; i < 4; i++) scores[t * 4 + i] = g.scores[i];
                    rankings(g.scores, nullptr, ranks + t * 4);
                    steps[t] = (int32_t)n_steps;
                }
            } catch (const std::exception& e) { errs[k] = e.what(); }
        };
        std::vector<std::thread> th;
        for (int k = 0; k < nt; k++) th.emplace_back(work, k);
        for (auto& t : th) t.join();
        for (auto& e : errs) if (!e.empty()) throw OrcError(e);
        if (used.load() != n_replay)
            throw OrcError("replay: " + std::to_string(n_replay - used.load()) + " recorded decisions were never requested");
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
// the same with a per-seat enable_quick_eval (uint8 [n_tables, 4]): challenger and champion may differ (agent/mortal.rs:54-74)
int orc_run_replay3(int n_tables, const uint64_t* nonces, const uint64_t* keys, int shuffle_kind, const uint8_t* qe_flags,
                    const int64_t* replay, int64_t n_replay, const int64_t* mask_bits, int64_t max_steps, int n_threads,
                    int32_t* scores, uint8_t* ranks, int32_t* steps) {
    g_replay_qe_flags = qe_flags;
    const int rc = orc_run_replay2(n_tables, nonces, keys, shuffle_kind, 1, replay, n_replay, mask_bits, max_steps, n_threads, scores, ranks, steps);
    g_replay_qe_flags = nullptr;
    return rc;
}
int orc_run_replay(int n_tables, const uint64_t* nonces, const uint64_t* keys, int shuffle_kind, int enable_quick_eval,
                   const int64_t* replay, int64_t n_replay, int32_t* scores, uint8_t* ranks, int32_t* steps) {
    return orc_run_replay2(n_tables, nonces, keys, shuffle_kind, enable_quick_eval, replay, n_replay, nullptr, 0, 1, scores, ranks, steps);
}

uint64_t orc_policy_hash(uint64_t nonce, uint64_t key, uint64_t table, uint64_t step_idx, uint32_t seat, uint32_t kan) {
    return policy_hash(nonce, key, table, step_idx, seat, kan);
}

// ---------------- dataset::GameplayLoader (dataset/gameplay.rs:247-449), SURVEY.md §8f N3 ----------------
// Replays one game's events from one player's point of view and records, at every decision the log shows the player
// making, the observation, legal mask and the label derived from the following events. Not restated: the oracle
// (invisible) observation, Grp and tile augmentation.
// Outputs are caller-allocated for `max_moves` entries; returns the number of moves (negative on error).
// dataset/invisible.rs Invisible (trust_seed branch: 35-66) + Invisible::encode (150-231): the hidden tiles of one kyoku,
// early -> late, regenerated from the game's seed
struct OrcInvisible {
    std::vector<u8> yama, rinshan, dora_indicators, ura_indicators;
    void encode(const PlayerState opp[3], size_t yama_idx, size_t rinshan_idx, int version, float* out) const {
        const int rows = oracle_obs_rows(version);
        for (int i = 0; i < rows * 34; i++) out[i] = 0.f;
        auto assign = [&](int r, int c, float v) { out[r * 34 + c] = v; };
        auto fill = [&](int r, float v) { for (int c = 0; c < 34; c++) out[r * 34 + c] = v; };
        int idx = 0;
        for (int k = 0; k < 3; k++) {
            const PlayerState& st = opp[k];
            for (int t = 0; t < 34; t++) for (int c = 0; c < st.tehai[t]; c++) assign(idx + c, t, 1.f);
            idx += 4;
            for (int i = 0; i < 3; i++) if (st.akas_in_hand[i]) fill(idx + i, 1.f);
            idx += 3;
            const int n = st.shanten;
            if (version == 1) { for (int i = 0; i < n; i++) fill(idx + i, 1.f); idx += 6; }
            else { fill(idx + n, 1.f); idx += 7; fill(idx, (float)n / 6.f); idx += 1; }
            for (int t = 0; t < 34; t++) if (st.waits[t]) assign(idx, t, 1.f);
            idx += 1;
            if (st.at_furiten) fill(idx, 1.f);
            idx += 1;
        }
        auto encode_tile = [&](int r, u8 tile) { assign(r, deaka(tile), 1.f); if (is_aka(tile)) fill(r + 1, 1.f); };
        for (size_t i = yama_idx; i < yama.size(); i++) { encode_tile(idx, yama[i]); idx += 2; }
        idx += ((int)yama_idx - 1) * 2;
        for (size_t i = rinshan_idx; i < rinshan.size(); i++) { encode_tile(idx, rinshan[i]); idx += 2; }
        idx += (int)rinshan_idx * 2;
        for (u8 t : dora_indicators) { encode_tile(idx, t); idx += 2; }
        for (u8 t : ura_indicators) { encode_tile(idx, t); idx += 2; }
        if (idx != rows) throw OrcError("Invisible::encode: row cursor mismatch");
    }
};

static int gameplay_load_impl(const orc_event* evs, int n_events, int player_id, int version, int always_include_kan_select, int sp_mode,
                              int max_moves, float* obs, uint8_t* masks, int64_t* actions, uint8_t* at_kyoku, uint8_t* apply_gamma,
                              uint8_t* at_turns, int8_t* shantens, bool oracle, uint64_t seed_nonce, uint64_t seed_key, int shuffle_kind,
                              float* inv_out, const uint8_t* walls);

int orc_gameplay_load(const orc_event* evs, int n_events, int player_id, int version, int always_include_kan_select, int sp_mode,
                      int max_moves, float* obs /*[max_moves, rows, 34] or null*/, uint8_t* masks /*[max_moves, 46]*/,
                      int64_t* actions, uint8_t* at_kyoku, uint8_t* apply_gamma, uint8_t* at_turns, int8_t* shantens) {
    return gameplay_load_impl(evs, n_events, player_id, version, always_include_kan_select, sp_mode, max_moves, obs, masks, actions,
                              at_kyoku, apply_gamma, at_turns, shantens, false, 0, 0, 0, nullptr, nullptr);
}
// the same with `oracle = true` (gameplay.rs:164, 308-331, 433-441): inv_out [max_moves, oracle_rows, 34]. The hidden tiles of
// kyoku q come from walls[q][136] (board.rs:109-122 layout) when given — what Invisible::new reconstructs from the log plus its
// random filler — else they are regenerated from the seed (`trust_seed`).
int orc_gameplay_load_oracle(const orc_event* evs, int n_events, int player_id, int version, int always_include_kan_select, int sp_mode,
                             int max_moves, float* obs, uint8_t* masks, int64_t* actions, uint8_t* at_kyoku, uint8_t* apply_gamma,
                             uint8_t* at_turns, int8_t* shantens, uint64_t seed_nonce, uint64_t seed_key, int shuffle_kind, float* inv_out,
                             const uint8_t* walls) {
    return gameplay_load_impl(evs, n_events, player_id, version, always_include_kan_select, sp_mode, max_moves, obs, masks, actions,
                              at_kyoku, apply_gamma, at_turns, shantens, true, seed_nonce, seed_key, shuffle_kind, inv_out, walls);
}

static int gameplay_load_impl(const orc_event* evs, int n_events, int player_id, int version, int always_include_kan_select, int sp_mode,
                              int max_moves, float* obs, uint8_t* masks, int64_t* actions, uint8_t* at_kyoku, uint8_t* apply_gamma,
                              uint8_t* at_turns, int8_t* shantens, bool oracle, uint64_t seed_nonce, uint64_t seed_key, int shuffle_kind,
                              float* inv_out, const uint8_t* walls) {
    try {
        PlayerState state((u8)player_id);
        const int rows = obs_rows(version);
        int kyoku_idx = 0, n = 0;
        std::vector<Event> ev(n_events);
        for (int i = 0; i < n_events; i++) ev[i] = from_c(evs[i]);
        // oracle: the three other seats' states (gameplay.rs:258-266) and the hidden tiles of every kyoku (invisible.rs:35-66)
        PlayerState opp[3] = {PlayerState((u8)((player_id + 1) % 4)), PlayerState((u8)((player_id + 2) % 4)), PlayerState((u8)((player_id + 3) % 4))};
        std::vector<OrcInvisible> invisibles;
        bool from_rinshan = false;
        size_t yama_idx = 0, rinshan_idx = 0;
        const int orows = oracle ? oracle_obs_rows(version) : 0;
        if (oracle)
            for (const Event& e : ev)
                if (e.type == EV_START_KYOKU) {
                    u8 seq[136];
                    if (walls) memcpy(seq, walls + invisibles.size() * 136, 136);
                    else make_wall(seed_nonce, seed_key, (u8)(4 * (e.bakaze - T_E) + e.kyoku - 1), e.honba, shuffle_kind, seq);
                    OrcInvisible iv;
                    for (int i = 135; i >= 66; i--) iv.yama.push_back(seq[i]);
                    for (int i = 55; i >= 52; i--) iv.rinshan.push_back(seq[i]);
                    for (int i = 60; i >= 56; i--) iv.dora_indicators.push_back(seq[i]);
                    for (int i = 61; i < 66; i++) iv.ura_indicators.push_back(seq[i]);
                    invisibles.push_back(std::move(iv));
                }
        auto add_entry = [&](bool at_kan_select, int label) {  // gameplay.rs:425-447
            if (n >= max_moves) throw OrcError("orc_gameplay_load: max_moves too small");
            std::vector<float> tmp;
            float* o = obs ? obs + (size_t)n * rows * 34 : nullptr;
            if (!o) { tmp.resize((size_t)rows * 34); o = tmp.data(); }
            state.encode_obs(version, at_kan_select, o, masks + (size_t)n * 46, sp_mode);
            actions[n] = label;
            at_kyoku[n] = (uint8_t)kyoku_idx;
            apply_gamma[n] = label <= 37;
            at_turns[n] = state.at_turn;
            shantens[n] = state.shanten;
            if (oracle && inv_out) invisibles.at(kyoku_idx).encode(opp, yama_idx, rinshan_idx, version, inv_out + (size_t)n * orows * 34);
            n++;
        };
        // gameplay.rs:279-283: windows of 4 events
        for (int w = 0; w + 4 <= n_events; w++) {
            const Event& cur = ev[w];
            const Event& next = (ev[w + 1].type == EV_REACH_ACCEPTED || ev[w + 1].type == EV_DORA) ? ev[w + 2] : ev[w + 1];
            if (cur.type == EV_END_KYOKU) kyoku_idx += 1;
            if (oracle) {  // gameplay.rs:308-331
                if (cur.type == EV_END_KYOKU) { from_rinshan = false; yama_idx = 0; rinshan_idx = 0; }
                else if (cur.type == EV_TSUMO) { if (from_rinshan) { rinshan_idx++; from_rinshan = false; } else yama_idx++; }
                else if (cur.type == EV_ANKAN || cur.type == EV_KAKAN || cur.type == EV_DAIMINKAN) from_rinshan = true;
                for (auto& s : opp) s.update(cur);
            }
            const ActionCandidate cans = state.update(cur);
            if (!cans.can_act()) continue;
            int label = -1, kan_select = -1;
            switch (next.type) {
                case EV_DAHAI: label = next.pai; break;
                case EV_REACH: label = 37; break;
                case EV_CHI:
                    if (next.actor == player_id) {
                        const u8 a = deaka(next.consumed[0]), b = deaka(next.consumed[1]), t = deaka(next.pai);
                        label = t < std::min(a, b) ? 38 : (t < std::max(a, b) ? 39 : 40);  // chi_type.rs:10-25
                    }
                    break;
                case EV_PON: if (next.actor == player_id) label = 41; break;
                case EV_DAIMINKAN:
                    if (next.actor == player_id) { if (always_include_kan_select) kan_select = deaka(next.pai); label = 42; }
                    break;
                case EV_KAKAN:
                    if (always_include_kan_select || state.kakan_candidates.size() > 1) kan_select = deaka(next.pai);
                    label = 42;
                    break;
                case EV_ANKAN:
                    if (always_include_kan_select || state.ankan_candidates.size() > 1) kan_select = deaka(next.consumed[0]);
                    label = 42;
                    break;
                case EV_RYUKYOKU: if (cans.can_ryukyoku) label = 44; break;
                default: break;
            }
            // the reference's match arms with guards fall through to the catch-all when the guard fails (gameplay.rs:349-416)
            const bool guarded_miss = label < 0 && next.type != EV_DAHAI && next.type != EV_REACH && next.type != EV_KAKAN &&
                                      next.type != EV_ANKAN;
            if (guarded_miss) {
                const bool has_any_ron = ev[w + 1].type == EV_HORA;
                if (has_any_ron) {
                    for (int k = w + 1; k < w + 4; k++) {
                        if (ev[k].type == EV_END_KYOKU) break;
                        if (ev[k].type == EV_HORA && ev[k].actor == player_id) { label = 43; break; }
                    }
                }
                if (label < 0) {
                    if ((cans.can_chi() && next.type == EV_TSUMO) ||
                        ((cans.can_pon || cans.can_daiminkan || cans.can_ron_agari) && !has_any_ron))
                        label = 45;
                }
            }
            if (label >= 0) {
                add_entry(false, label);
                if (kan_select >= 0) add_entry(true, kan_select);
            }
        }
        return n;
    } catch (const std::exception& e) { return fail(e); }
}

}  // extern "C"
