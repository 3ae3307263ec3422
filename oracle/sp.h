// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi algo/sp/{calc,state,candidate,tile}.rs and
// state/agent_helper.rs:509-593 (single_player_tables).
#pragma once
#include "state.h"

namespace orc {

struct RequiredTile { u8 tile; u8 count; };

// sp/candidate.rs:9-25
struct SpCandidate {
    u8 tile = T_UNK;
    std::vector<float> tenpai_probs, win_probs, exp_values;
    std::vector<RequiredTile> required_tiles;
    u8 num_required_tiles = 0;
    bool shanten_down = false;
};

enum SpColumn { SPCOL_EV, SPCOL_WIN_PROB, SPCOL_TENPAI_PROB, SPCOL_NOT_SHANTEN_DOWN, SPCOL_NUM_REQUIRED, SPCOL_DISCARD_PRIORITY };
// sp/candidate.rs:73-107
int sp_candidate_cmp(const SpCandidate& l, const SpCandidate& r, SpColumn by);

// sp/state.rs:24-33
struct SpInitState {
    u8 tehai[34];
    bool akas_in_hand[3];
    u8 tiles_seen[34];
    bool akas_seen[3];
};

// sp/calc.rs:36-62
struct SpCalculator {
    u8 tehai_len_div3 = 4;
    const u8* chis = nullptr; int n_chis = 0;
    const u8* pons = nullptr; int n_pons = 0;
    const u8* minkans = nullptr; int n_minkans = 0;
    const u8* ankans = nullptr; int n_ankans = 0;
    u8 bakaze = T_E, jikaze = T_E;
    bool is_menzen = true;
    u8 num_doras_in_fuuro = 0;
    const u8* dora_indicators = nullptr; int n_dora_indicators = 0;
    bool calc_double_riichi = false, calc_haitei = false, prefer_riichi = true, sort_result = true;
    bool maximize_win_prob = false, calc_tegawari = false, calc_shanten_down = false;

    // sp/calc.rs:84-134
    std::vector<SpCandidate> calc(const SpInitState& init, bool can_discard, u8 tsumos_left, i8 cur_shanten) const;
};

// agent_helper.rs:509-593. Returns false where the reference returns Err.
bool single_player_tables(const PlayerState& st, std::vector<SpCandidate>& out);

}  // namespace orc
