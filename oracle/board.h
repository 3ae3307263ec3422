// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi arena/{board,game,result,one_vs_three}.rs and agent/mortal.rs
// (set_scene quick-eval / kan-select, get_reaction action decode).
#pragma once
#include "state.h"

#include <functional>

namespace orc {

// ---- wall: arena/board.rs:99-123, 786-824; third-party sha3 0.10.8 / rand_chacha 0.9.0 /
// rand 0.9.1 (Cargo.lock:1236,1052,1042) restated from their published algorithms.
void sha3_256(const u8* data, size_t len, u8* out32);
struct ChaCha12 {
    u32 key[8]; u64 counter = 0; u32 buf[16]; int pos = 16;
    explicit ChaCha12(const u8* seed32);
    u32 next_u32();
};
enum ShuffleKind { SHUFFLE_RAND09 = 0, SHUFFLE_RAND08 = 1 };
// seq[136]: shuffled wall (board.rs:109-110)
void make_wall(u64 nonce, u64 key, u8 kyoku, u8 honba, int shuffle_kind, u8* seq136);

// arena/board.rs:30-48
struct Board {
    u8 kyoku = 0, honba = 0, kyotaku = 0;
    i32 scores[4] = {25000, 25000, 25000, 25000};
    u8 haipai[4][13];
    std::vector<u8> yama, rinshan, dora_indicators, ura_indicators;
    void init_from_wall(const u8* seq136);  // board.rs:111-122
};

enum Poll { POLL_INGAME, POLL_END };

// arena/result.rs:8-17
struct KyokuResult {
    u8 kyoku; bool can_renchan, has_hora, has_abortive_ryukyoku; u8 kyotaku_left; i32 scores[4];
};

// arena/board.rs:52-85
struct BoardState {
    Board board;
    u8 oya = 0;
    PlayerState player_states[4];
    bool can_renchan = false, has_hora = false, has_abortive_ryukyoku = false;
    i32 kyoku_deltas[4] = {0, 0, 0, 0};
    u8 tiles_left = 70;
    u8 tsumo_actor = 0;
    bool deal_from_rinshan = false, need_new_dora_at_discard = false, need_new_dora_at_tsumo = false;
    int riichi_to_be_accepted = -1;
    bool can_nagashi_mangan[4] = {true, true, true, true};
    bool can_four_wind = true;
    int four_wind_tile = -1;
    u8 accepted_riichis = 0, kans = 0;
    bool check_four_kan = false;
    int paos[4] = {-1, -1, -1, -1};
    std::vector<Event> log;
    std::vector<u8> dora_indicators_full;  // board.rs:84: all five indicators as dealt, for encode_oracle_obs

    explicit BoardState(const Board& b);  // board.rs:125-137 into_state
    // board.rs:680-782: the invisible (oracle) observation from `perspective`: [oracle_obs_rows(version)][34] f32, zero-filled here
    void encode_oracle_obs(u8 perspective, int version, float* out) const;
    Poll poll(const Event reactions[4]);  // board.rs:141-161
    KyokuResult end() const;              // board.rs:172-182

private:
    Poll step(const Event reactions[4]);  // board.rs:511-678
    void broadcast(const Event& ev);
    void haipai();
    void exhaustive_ryukyoku();
    void update_nagashi_mangan_and_four_wind(const Event& ev);
    bool check_four_wind(u8 pai);
    void check_riichi_accepted();
    void add_new_dora();
    void handle_hora(u8 single_actor, u8 single_target, const Event reactions[4]);
    void update_paos(const Event& ev);
    void abortive_ryukyoku();
};

// ---- agent side: agent/mortal.rs ----
struct AgentConfig {
    bool enable_quick_eval = true;             // engine.py:18
    bool enable_rule_based_agari_guard = false;
    int version = 4;
};
// A decision request: one row the policy must answer (mortal.rs:244-287)
struct Scene {
    int table = 0;     // game index in the batch
    u8 seat = 0;       // absolute seat
    bool is_kan_select = false;
    u64 step_idx = 0;  // table-step counter of this table (0-based)
    const PlayerState* state = nullptr;
    const BoardState* board = nullptr;  // the full-information board (what an oracle agent's invisible_obs is encoded from)
};
int oracle_obs_rows(int version);  // consts.rs:30-38 oracle_obs_shape(version).0
// policy: (scene, mask[46]) -> action id. q-values for the agari guard are optional.
typedef std::function<int(const Scene&, const u8* mask46, float* q46_or_null)> PolicyFn;

// mortal.rs:338-573 — action id -> Event (throws OrcError on failed checks)
Event decode_action(const PlayerState& st, u8 actor, int action, int kan_select_action /* -1 if none */);
// mask only (no planes): obs_repr.rs mask writes
void legal_mask(const PlayerState& st, bool at_kan_select, u8* mask46);

// arena/game.rs:28-218
struct Game {
    u8 length = 8;
    u64 seed_nonce = 0, seed_key = 0;
    int shuffle_kind = 0;
    int table = 0;
    Event last_reactions[4];
    BoardState* board = nullptr;
    u8 kyoku = 0, honba = 0, kyotaku = 0;
    i32 scores[4] = {25000, 25000, 25000, 25000};
    std::vector<std::vector<Event>> game_log;
    bool kyoku_started = false, ended = false, in_renchan = false;
    u64 step_idx = 0;
    // injected walls for tests (board.rs:20-21: fields are pub so callers may set the yama)
    std::function<bool(u8 kyoku, u8 honba, u8* seq136)> wall_override;

    ~Game() { delete board; }
    void poll();                                            // game.rs:59-178
    // returns true if the game has ended (game.rs:180-198), else collects reactions (200-217)
    bool commit(const AgentConfig cfgs[4], const PolicyFn pols[4], std::vector<int>* action_trace);
};

// built-in test policies (shared definition with mortal_b200/csrc/policy_test.cu)
u64 splitmix64(u64 x);
u64 policy_hash(u64 nonce, u64 key, u64 table, u64 step_idx, u32 seat, u32 kan);
// kind 0: uniform over mask; kind 1: agari-first / shanten-greedy (see board.cc)
int test_policy(int kind, const Scene& sc, u64 nonce, u64 key, const u8* mask46);

}  // namespace orc
