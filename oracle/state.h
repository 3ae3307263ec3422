// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi mjai/event.rs (Event), state/{player_state,item,action,update,
// agent_helper,obs_repr}.rs.
#pragma once
#include "mj.h"

namespace orc {

// mjai/event.rs:20-120 — numeric form (no JSON on the hot path)
enum EvType : u8 {
    EV_NONE = 0, EV_START_GAME, EV_START_KYOKU, EV_TSUMO, EV_DAHAI, EV_CHI, EV_PON, EV_DAIMINKAN,
    EV_KAKAN, EV_ANKAN, EV_DORA, EV_REACH, EV_REACH_ACCEPTED, EV_HORA, EV_RYUKYOKU, EV_END_KYOKU,
    EV_END_GAME,
};

struct Event {
    u8 type = EV_NONE;
    u8 actor = 0, target = 0;
    u8 pai = T_UNK;  // also dora_marker for EV_DORA / EV_START_KYOKU
    bool tsumogiri = false;
    u8 consumed[4] = {T_UNK, T_UNK, T_UNK, T_UNK};
    // start_kyoku
    u8 bakaze = T_E, kyoku = 1, honba = 0, kyotaku = 0, oya = 0;
    i32 scores[4] = {0, 0, 0, 0};
    u8 tehais[4][13] = {};
    // hora / ryukyoku
    bool has_deltas = false;
    i32 deltas[4] = {0, 0, 0, 0};
    u8 ura_markers[5] = {}; int n_ura = 0;

    // event.rs:158-174
    bool has_actor() const {
        switch (type) {
            case EV_TSUMO: case EV_DAHAI: case EV_CHI: case EV_PON: case EV_DAIMINKAN: case EV_KAKAN:
            case EV_ANKAN: case EV_REACH: case EV_REACH_ACCEPTED: case EV_HORA: return true;
            default: return false;
        }
    }
    // event.rs:178-183
    bool is_in_game_announce() const { return type == EV_REACH_ACCEPTED || type == EV_DORA || type == EV_HORA; }
};

// state/item.rs:7-27
struct Sutehai { u8 tile = T_UNK; bool is_dora = false, is_tedashi = false, is_riichi = false; };
struct ChiPon { u8 consumed[2] = {T_UNK, T_UNK}; u8 target_tile = T_UNK; };
struct KawaItem {
    bool has_chi_pon = false; ChiPon chi_pon;
    u8 kan[4] = {}; int n_kan = 0;
    Sutehai sutehai;
};
struct KawaSlot { bool some = false; KawaItem item; };
struct OptSutehai { bool some = false; Sutehai s; };

// state/action.rs:11-89
struct ActionCandidate {
    bool can_discard = false, can_chi_low = false, can_chi_mid = false, can_chi_high = false, can_pon = false,
         can_daiminkan = false, can_kakan = false, can_ankan = false, can_riichi = false,
         can_tsumo_agari = false, can_ron_agari = false, can_ryukyoku = false;
    u8 target_actor = 0;
    bool can_chi() const { return can_chi_low || can_chi_mid || can_chi_high; }
    bool can_kan() const { return can_daiminkan || can_kakan || can_ankan; }
    bool can_agari() const { return can_tsumo_agari || can_ron_agari; }
    bool can_pass() const { return can_chi() || can_pon || can_daiminkan || can_ron_agari; }
    bool can_act() const {
        return can_discard || can_chi() || can_pon || can_kan() || can_riichi || can_agari() || can_ryukyoku;
    }
};

enum MoveType { MOVE_TSUMO, MOVE_DISCARD, MOVE_FUURO_CONSUME };

// state/player_state.rs:24-140
struct PlayerState {
    u8 player_id = 0;
    u8 tehai[34] = {};
    bool waits[34] = {};
    u8 dora_factor[34] = {};
    u8 tiles_seen[34] = {};
    bool akas_seen[3] = {};
    bool keep_shanten_discards[34] = {};
    bool next_shanten_discards[34] = {};
    bool forbidden_tiles[34] = {};
    bool discarded_tiles[34] = {};
    u8 bakaze = T_UNK, jikaze = T_UNK;
    u8 kyoku = 0, honba = 0, kyotaku = 0;
    i32 scores[4] = {};
    u8 rank = 0;
    u8 oya = 0;
    bool is_all_last = false;
    std::vector<u8> dora_indicators;
    std::vector<KawaSlot> kawa[4];
    OptSutehai last_tedashis[4];
    OptSutehai riichi_sutehais[4];
    std::vector<u8> kawa_overview[4];
    std::vector<std::vector<u8>> fuuro_overview[4];
    std::vector<u8> ankan_overview[4];
    bool riichi_declared[4] = {};
    bool riichi_accepted[4] = {};
    u8 at_turn = 0;
    u8 tiles_left = 0;
    std::vector<u8> intermediate_kan;
    bool has_intermediate_chi_pon = false; ChiPon intermediate_chi_pon;
    i8 shanten = 0;
    bool has_last_self_tsumo = false; u8 last_self_tsumo = T_UNK;
    bool has_last_kawa_tile = false; u8 last_kawa_tile = T_UNK;
    ActionCandidate last_cans;
    std::vector<u8> ankan_candidates, kakan_candidates;
    bool chankan_chance = false;
    bool can_w_riichi = false, is_w_riichi = false, at_rinshan = false, at_ippatsu = false, at_furiten = false;
    bool to_mark_same_cycle_furiten = false;
    u8 kans_on_board = 0;
    bool is_menzen = false;
    std::vector<u8> chis, pons, minkans, ankans;
    u8 doras_owned[4] = {};
    u8 doras_seen = 0;
    bool akas_in_hand[3] = {};
    u8 tehai_len_div3 = 0;
    bool has_next_shanten_discard = false;

    explicit PlayerState(u8 pid = 0) : player_id(pid) {}

    // update.rs
    ActionCandidate update(const Event& ev, bool keep_cans_on_announce = false);
    int rel(u8 actor) const { return (actor + 4 - player_id) % 4; }
    void witness_tile(u8 tile);
    void move_tile(u8 tile, MoveType mt);
    void add_dora_indicator(u8 tile);
    void pad_kawa_for_pon_or_daiminkan(u8 abs_actor, u8 abs_target);
    void pad_kawa_at_start();
    void set_can_chi_from_tile(u8 tile);
    void update_shanten();
    void update_shanten_discards();
    void update_waits_and_furiten();
    void update_doras_owned(int actor_rel, u8 tile);
    void update_rank();
    u8 get_rank(const i32* scores_rel) const;

    // action.rs:93-227 — throws OrcError on invalid reaction
    void validate_reaction(const Event& action) const;

    // agent_helper.rs
    int kans_count() const { return (int)(minkans.size() + ankans.size()); }
    void discard_candidates_aka(bool* out37) const;
    void discard_candidates_with_unconditional_tenpai_aka(bool* out37) const;
    void discard_candidates_with_unconditional_tenpai(bool* out34) const;
    u8 yaokyuu_kind_count() const;
    bool rule_based_agari() const;
    bool rule_based_agari_slow(bool is_ron, int target_rel) const;
    Point agari_points(bool is_ron, const u8* ura, int n_ura) const;  // throws if cannot agari
    i8 real_time_shanten() const;
    bool is_oya() const { return oya == 0; }

    // obs_repr.rs:126-630 — obs: [rows(version)*34] f32 (zeroed here), mask: [46].
    // sp_mode: 0 = SP rows left zero except max-EV fallback semantics documented in obs.cc,
    //          1 = full single_player_tables (agent_helper.rs:509-593)
    void encode_obs(int version, bool at_kan_select, float* obs, u8* mask46, int sp_mode = 1) const;

private:
    void start_kyoku(const Event& ev);
    void tsumo(u8 actor, u8 pai);
    void dahai(u8 actor, u8 pai, bool tsumogiri);
    void chi(u8 actor, u8 pai, const u8* consumed);
    void pon(u8 actor, u8 target, u8 pai, const u8* consumed);
    void daiminkan(u8 actor, u8 target, u8 pai, const u8* consumed);
    void kakan(u8 actor, u8 pai);
    void ankan(u8 actor, const u8* consumed);
    void reach(u8 actor);
    void reach_accepted(u8 actor);
    void ensure_tiles_in_hand(const u8* tiles, int n) const;
    AgariCalc make_agari_calc(const u8* tehai34, u8 winning_tile, bool is_ron) const;
};

int obs_rows(int version);  // consts.rs:20-28

}  // namespace orc
