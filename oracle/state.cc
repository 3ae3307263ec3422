// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi state/update.rs, state/action.rs, state/agent_helper.rs (minus SP).
#include "state.h"

#include <algorithm>

namespace orc {

static inline bool vec_contains(const std::vector<u8>& v, u8 x) {
    return std::find(v.begin(), v.end(), x) != v.end();
}

AgariCalc PlayerState::make_agari_calc(const u8* tehai34, u8 winning_tile, bool is_ron) const {
    AgariCalc c;
    c.tehai = tehai34;
    c.is_menzen = is_menzen;
    c.chis = chis.data(); c.n_chis = (int)chis.size();
    c.pons = pons.data(); c.n_pons = (int)pons.size();
    c.minkans = minkans.data(); c.n_minkans = (int)minkans.size();
    c.ankans = ankans.data(); c.n_ankans = (int)ankans.size();
    c.bakaze = bakaze; c.jikaze = jikaze;
    c.winning_tile = winning_tile;
    c.is_ron = is_ron;
    return c;
}

// update.rs:24-122
ActionCandidate PlayerState::update(const Event& ev, bool keep_cans_on_announce) {
    if (!keep_cans_on_announce || !ev.is_in_game_announce()) {
        last_cans = ActionCandidate{};
        last_cans.target_actor = ev.has_actor() ? ev.actor : player_id;
        ankan_candidates.clear();
        kakan_candidates.clear();
    }
    if (to_mark_same_cycle_furiten) { to_mark_same_cycle_furiten = false; at_furiten = true; }
    if (chankan_chance) { chankan_chance = false; at_ippatsu = false; }

    switch (ev.type) {
        case EV_START_KYOKU: start_kyoku(ev); break;
        case EV_TSUMO: tsumo(ev.actor, ev.pai); break;
        case EV_DAHAI: dahai(ev.actor, ev.pai, ev.tsumogiri); break;
        case EV_CHI: chi(ev.actor, ev.pai, ev.consumed); break;
        case EV_PON: pon(ev.actor, ev.target, ev.pai, ev.consumed); break;
        case EV_DAIMINKAN: daiminkan(ev.actor, ev.target, ev.pai, ev.consumed); break;
        case EV_KAKAN: kakan(ev.actor, ev.pai); break;
        case EV_ANKAN: ankan(ev.actor, ev.consumed); break;
        case EV_DORA: add_dora_indicator(ev.pai); break;
        case EV_REACH: reach(ev.actor); break;
        case EV_REACH_ACCEPTED: reach_accepted(ev.actor); break;
        default: break;
    }
    return last_cans;
}

// update.rs:125-217
void PlayerState::start_kyoku(const Event& ev) {
    memset(tehai, 0, sizeof tehai);
    memset(waits, 0, sizeof waits);
    memset(dora_factor, 0, sizeof dora_factor);
    memset(tiles_seen, 0, sizeof tiles_seen);
    memset(akas_seen, 0, sizeof akas_seen);
    memset(keep_shanten_discards, 0, sizeof keep_shanten_discards);
    memset(next_shanten_discards, 0, sizeof next_shanten_discards);
    memset(forbidden_tiles, 0, sizeof forbidden_tiles);
    memset(discarded_tiles, 0, sizeof discarded_tiles);

    bakaze = ev.bakaze;
    honba = ev.honba;
    kyotaku = ev.kyotaku;
    oya = (u8)rel(ev.oya);
    jikaze = T_E + (4 - oya) % 4;
    kyoku = ev.kyoku - 1;
    is_all_last = bakaze == T_E ? false : bakaze == T_S ? kyoku == 3 : true;

    for (int i = 0; i < 4; i++) scores[i] = ev.scores[(i + player_id) % 4];  // rotate_left(player_id)

    dora_indicators.clear();
    memset(doras_owned, 0, sizeof doras_owned);
    doras_seen = 0;
    memset(akas_in_hand, 0, sizeof akas_in_hand);

    ankan_candidates.clear();
    kakan_candidates.clear();
    chankan_chance = false;

    at_ippatsu = false;
    at_rinshan = false;
    at_furiten = false;
    to_mark_same_cycle_furiten = false;

    is_menzen = true;
    can_w_riichi = true;
    is_w_riichi = false;
    chis.clear(); pons.clear(); minkans.clear(); ankans.clear();

    kans_on_board = 0;
    tehai_len_div3 = 4;
    has_next_shanten_discard = false;
    tiles_left = 70;
    at_turn = 0;

    for (int i = 0; i < 4; i++) {
        kawa[i].clear();
        last_tedashis[i] = OptSutehai{};
        kawa_overview[i].clear();
        fuuro_overview[i].clear();
        ankan_overview[i].clear();
        riichi_declared[i] = false;
        riichi_accepted[i] = false;
        riichi_sutehais[i] = OptSutehai{};
    }
    intermediate_kan.clear();
    has_intermediate_chi_pon = false;

    has_last_self_tsumo = false;
    has_last_kawa_tile = false;

    update_rank();
    add_dora_indicator(ev.pai);
    for (int i = 0; i < 13; i++) {
        u8 t = ev.tehais[player_id][i];
        witness_tile(t);
        move_tile(t, MOVE_TSUMO);
    }
    update_shanten();
    update_waits_and_furiten();
    pad_kawa_at_start();
}

// update.rs:219-309
void PlayerState::tsumo(u8 actor, u8 pai) {
    ORC_ENSURE(tiles_left > 0, "rule violation: attempt to tsumo from exhausted yama");
    tiles_left -= 1;
    if (actor != player_id) return;
    at_turn += 1;

    last_cans.can_discard = true;
    has_last_self_tsumo = true; last_self_tsumo = pai;
    witness_tile(pai);
    move_tile(pai, MOVE_TSUMO);

    if (can_w_riichi) last_cans.can_ryukyoku = yaokyuu_kind_count() >= 9;

    if (!riichi_accepted[0]) update_shanten_discards();

    u8 pid = deaka(pai);
    if (waits[pid]) {
        if (is_menzen || riichi_accepted[0] || tiles_left == 0 || at_rinshan || can_w_riichi) {
            last_cans.can_tsumo_agari = true;
        } else {
            AgariCalc c = make_agari_calc(tehai, pid, false);
            last_cans.can_tsumo_agari = c.has_yaku();
        }
    }

    if (tiles_left == 0) return;

    if (riichi_accepted[0]) {
        if (kans_on_board < 4) {
            last_cans.can_ankan = check_ankan_after_riichi(tehai, tehai_len_div3, pai, false);
            if (last_cans.can_ankan) ankan_candidates.push_back(pid);
        }
        return;
    }

    if (kans_on_board < 4) {
        for (int tid = 0; tid < 34; tid++) {
            u8 count = tehai[tid];
            if (count == 0) continue;
            if (count == 4) {
                last_cans.can_ankan = true;
                ankan_candidates.push_back((u8)tid);
            } else if (vec_contains(pons, (u8)tid)) {
                last_cans.can_kakan = true;
                kakan_candidates.push_back((u8)tid);
            }
        }
    }

    last_cans.can_riichi = is_menzen && tiles_left >= 4 && scores[0] >= 1000 &&
                           (shanten == 0 || (shanten == 1 && has_next_shanten_discard));
}

// update.rs:311-427
void PlayerState::dahai(u8 actor, u8 pai, bool tsumogiri) {
    int actor_rel = rel(actor);
    if (actor_rel == 0) move_tile(pai, MOVE_DISCARD);
    else witness_tile(pai);

    u8 pid = deaka(pai);
    bool is_riichi = riichi_declared[actor_rel] && !riichi_accepted[actor_rel];
    Sutehai sutehai;
    sutehai.tile = pai;
    sutehai.is_dora = dora_factor[pid] > 0;
    sutehai.is_tedashi = !tsumogiri;
    sutehai.is_riichi = is_riichi;
    KawaSlot slot;
    slot.some = true;
    slot.item.n_kan = (int)intermediate_kan.size();
    for (int i = 0; i < slot.item.n_kan; i++) slot.item.kan[i] = intermediate_kan[i];
    intermediate_kan.clear();
    slot.item.has_chi_pon = has_intermediate_chi_pon;
    slot.item.chi_pon = intermediate_chi_pon;
    has_intermediate_chi_pon = false;
    slot.item.sutehai = sutehai;
    kawa[actor_rel].push_back(slot);
    kawa_overview[actor_rel].push_back(pai);
    has_last_kawa_tile = true; last_kawa_tile = pai;

    if (!tsumogiri) { last_tedashis[actor_rel].some = true; last_tedashis[actor_rel].s = sutehai; }
    if (is_riichi) { riichi_sutehais[actor_rel].some = true; riichi_sutehais[actor_rel].s = sutehai; }

    if (actor_rel == 0) {
        memset(forbidden_tiles, 0, sizeof forbidden_tiles);
        at_rinshan = false;
        at_ippatsu = false;
        can_w_riichi = false;
        discarded_tiles[pid] = true;

        if (!riichi_accepted[0]) {
            if (next_shanten_discards[pid]) shanten -= 1;
            else if (!keep_shanten_discards[pid]) update_shanten();
            update_waits_and_furiten();
        } else if (!at_furiten && waits[pid]) {
            at_furiten = true;
        }
        return;
    }

    if (!at_furiten && waits[pid]) {
        if (riichi_accepted[0] || tiles_left == 0) {
            last_cans.can_ron_agari = true;
        } else {
            u8 th[34];
            memcpy(th, tehai, 34);
            th[pid] += 1;
            AgariCalc c = make_agari_calc(th, pid, true);
            last_cans.can_ron_agari = c.has_yaku();
        }
        if (last_cans.can_ron_agari) to_mark_same_cycle_furiten = true;
        else at_furiten = true;
    }

    if (riichi_accepted[0] || tiles_left == 0) return;

    if (actor_rel == 3 && !is_jihai(pai) && tehai_len_div3 > 0) set_can_chi_from_tile(pai);
    last_cans.can_pon = tehai[pid] >= 2;
    last_cans.can_daiminkan = kans_on_board < 4 && tehai[pid] == 3;
}

// update.rs:429-495
void PlayerState::chi(u8 actor, u8 pai, const u8* consumed) {
    int actor_rel = rel(actor);
    std::vector<u8> full_set = {consumed[0], consumed[1], pai};
    fuuro_overview[actor_rel].push_back(full_set);
    has_intermediate_chi_pon = true;
    intermediate_chi_pon.consumed[0] = consumed[0];
    intermediate_chi_pon.consumed[1] = consumed[1];
    intermediate_chi_pon.target_tile = pai;

    if (actor_rel != 0) {
        for (int i = 0; i < 2; i++) witness_tile(consumed[i]);
        for (u8 t : full_set) update_doras_owned(actor_rel, t);
        can_w_riichi = false;
        at_ippatsu = false;
        return;
    }

    last_cans.can_discard = true;
    is_menzen = false;
    tehai_len_div3 -= 1;
    has_last_self_tsumo = false;

    update_doras_owned(0, pai);
    for (int i = 0; i < 2; i++) move_tile(consumed[i], MOVE_FUURO_CONSUME);

    int a = deaka(consumed[0]), b = deaka(consumed[1]);
    int mn = std::min(a, b), mx = std::max(a, b);
    int tid = deaka(pai);
    chis.push_back((u8)std::min(mn, tid));

    if (tehai[tid] > 0) forbidden_tiles[tid] = true;
    if (tid < mn) {
        if (mx % 9 < 8) {
            int bigger = mx + 1;
            if (tehai[bigger] > 0) forbidden_tiles[bigger] = true;
        }
    } else if (tid > mx && mn % 9 > 0) {
        int smaller = mn - 1;
        if (tehai[smaller] > 0) forbidden_tiles[smaller] = true;
    }

    update_shanten();
    update_shanten_discards();
}

// update.rs:497-542
void PlayerState::pon(u8 actor, u8 target, u8 pai, const u8* consumed) {
    int actor_rel = rel(actor);
    std::vector<u8> full_set = {consumed[0], consumed[1], pai};
    fuuro_overview[actor_rel].push_back(full_set);
    has_intermediate_chi_pon = true;
    intermediate_chi_pon.consumed[0] = consumed[0];
    intermediate_chi_pon.consumed[1] = consumed[1];
    intermediate_chi_pon.target_tile = pai;
    pad_kawa_for_pon_or_daiminkan(actor, target);

    if (actor_rel != 0) {
        for (int i = 0; i < 2; i++) witness_tile(consumed[i]);
        for (u8 t : full_set) update_doras_owned(actor_rel, t);
        can_w_riichi = false;
        at_ippatsu = false;
        return;
    }

    last_cans.can_discard = true;
    is_menzen = false;
    tehai_len_div3 -= 1;
    has_last_self_tsumo = false;

    update_doras_owned(0, pai);
    for (int i = 0; i < 2; i++) move_tile(consumed[i], MOVE_FUURO_CONSUME);
    u8 pid = deaka(pai);
    pons.push_back(pid);

    if (tehai[pid] > 0) forbidden_tiles[pid] = true;

    update_shanten();
    update_shanten_discards();
}

// update.rs:544-582
void PlayerState::daiminkan(u8 actor, u8 target, u8 pai, const u8* consumed) {
    int actor_rel = rel(actor);
    std::vector<u8> full_set = {consumed[0], consumed[1], consumed[2], pai};
    fuuro_overview[actor_rel].push_back(full_set);
    intermediate_kan.push_back(pai);
    pad_kawa_for_pon_or_daiminkan(actor, target);
    kans_on_board += 1;

    if (actor_rel != 0) {
        for (int i = 0; i < 3; i++) witness_tile(consumed[i]);
        for (u8 t : full_set) update_doras_owned(actor_rel, t);
        can_w_riichi = false;
        at_ippatsu = false;
        return;
    }

    at_rinshan = true;
    is_menzen = false;
    tehai_len_div3 -= 1;

    update_doras_owned(0, pai);
    for (int i = 0; i < 3; i++) move_tile(consumed[i], MOVE_FUURO_CONSUME);
    minkans.push_back(deaka(pai));

    update_shanten();
    update_waits_and_furiten();
}

// update.rs:584-628
void PlayerState::kakan(u8 actor, u8 pai) {
    int actor_rel = rel(actor);
    u8 pid = deaka(pai);
    for (auto& fuuro : fuuro_overview[actor_rel]) {
        if (deaka(fuuro[0]) == pid) {
            fuuro.push_back(pai);
            break;
        }
    }
    intermediate_kan.push_back(pai);
    kans_on_board += 1;

    if (actor_rel != 0) {
        witness_tile(pai);
        update_doras_owned(actor_rel, pai);
        has_last_kawa_tile = true; last_kawa_tile = pai;

        if (!at_furiten && waits[pid]) {
            last_cans.can_ron_agari = true;
            to_mark_same_cycle_furiten = true;
            chankan_chance = true;
        } else {
            at_ippatsu = false;
        }
        return;
    }

    at_rinshan = true;
    move_tile(pai, MOVE_FUURO_CONSUME);
    pons.erase(std::remove(pons.begin(), pons.end(), pid), pons.end());
    minkans.push_back(pid);

    if (next_shanten_discards[pid]) shanten -= 1;
    else if (!keep_shanten_discards[pid]) update_shanten();
    update_waits_and_furiten();
}

// update.rs:630-663
void PlayerState::ankan(u8 actor, const u8* consumed) {
    int actor_rel = rel(actor);
    u8 tile = deaka(consumed[0]);
    ankan_overview[actor_rel].push_back(tile);
    intermediate_kan.push_back(tile);
    kans_on_board += 1;

    can_w_riichi = false;
    at_ippatsu = false;

    if (actor_rel != 0) {
        for (int i = 0; i < 4; i++) {
            witness_tile(consumed[i]);
            update_doras_owned(actor_rel, consumed[i]);
        }
        return;
    }

    at_rinshan = true;
    tehai_len_div3 -= 1;
    for (int i = 0; i < 4; i++) move_tile(consumed[i], MOVE_FUURO_CONSUME);
    ankans.push_back(tile);

    if (!riichi_accepted[0]) {
        update_shanten();
        update_waits_and_furiten();
    }
}

// update.rs:665-675
void PlayerState::reach(u8 actor) {
    int actor_rel = rel(actor);
    riichi_declared[actor_rel] = true;
    if (actor_rel == 0) {
        is_w_riichi = can_w_riichi;
        last_cans.can_discard = true;
    }
}

// update.rs:677-686
void PlayerState::reach_accepted(u8 actor) {
    int actor_rel = rel(actor);
    riichi_accepted[actor_rel] = true;
    scores[actor_rel] -= 1000;
    kyotaku += 1;
    update_rank();
    if (actor_rel == 0) at_ippatsu = true;
}

// update.rs:695-727
void PlayerState::witness_tile(u8 tile) {
    ORC_ENSURE(tile < T_UNK, "rule violation: attempt to witness an unknown tile");
    u8 tid = deaka(tile);
    ORC_ENSURE(tiles_seen[tid] < 4, "rule violation: attempt to witness the fifth tile");
    tiles_seen[tid] += 1;
    doras_seen += dora_factor[tid];
    if (tile == T_5MR) { akas_seen[0] = true; doras_seen += 1; }
    else if (tile == T_5PR) { akas_seen[1] = true; doras_seen += 1; }
    else if (tile == T_5SR) { akas_seen[2] = true; doras_seen += 1; }
}

// update.rs:734-775
void PlayerState::move_tile(u8 tile, MoveType mt) {
    u8 tid = deaka(tile);
    switch (mt) {
        case MOVE_TSUMO:
            tehai[tid] += 1;
            doras_owned[0] += dora_factor[tid];
            break;
        case MOVE_DISCARD:
            ORC_ENSURE(tehai[tid] > 0, "rule violation: attempt to discard from void");
            tehai[tid] -= 1;
            doras_owned[0] -= dora_factor[tid];
            break;
        case MOVE_FUURO_CONSUME:
            ORC_ENSURE(tehai[tid] > 0, "rule violation: attempt to consume from void");
            tehai[tid] -= 1;
            break;
    }
    if (is_aka(tile)) {
        int aka_id = tile - T_5MR;
        switch (mt) {
            case MOVE_TSUMO: akas_in_hand[aka_id] = true; doras_owned[0] += 1; break;
            case MOVE_DISCARD: akas_in_hand[aka_id] = false; doras_owned[0] -= 1; break;
            case MOVE_FUURO_CONSUME: akas_in_hand[aka_id] = false; break;
        }
    }
}

// update.rs:780-808
void PlayerState::add_dora_indicator(u8 tile) {
    ORC_ENSURE(dora_indicators.size() < 5, "too many dora indicators");
    dora_indicators.push_back(tile);
    witness_tile(tile);
    u8 next = tile_next(tile);
    dora_factor[next] += 1;
    doras_owned[0] += tehai[next];
    for (int i = 0; i < 4; i++) {
        int cnt = 0;
        for (auto& f : fuuro_overview[i])
            for (u8 t : f) if (deaka(t) == next) cnt++;
        doras_owned[i] += (u8)cnt;
        if (vec_contains(ankan_overview[i], next)) doras_owned[i] += 4;
    }
    doras_seen += tiles_seen[next];
}

// update.rs:810-817
void PlayerState::pad_kawa_for_pon_or_daiminkan(u8 abs_actor, u8 abs_target) {
    u8 i = (abs_target + 1) % 4;
    while (i != abs_actor) {
        kawa[rel(i)].push_back(KawaSlot{});
        i = (i + 1) % 4;
    }
}

// update.rs:819-824
void PlayerState::pad_kawa_at_start() {
    for (int i = 0; i < oya && i < 4; i++) kawa[i].push_back(KawaSlot{});
}

// update.rs:826-868
void PlayerState::set_can_chi_from_tile(u8 tile) {
    last_cans.can_chi_low = last_cans.can_chi_mid = last_cans.can_chi_high = false;
    int tid = deaka(tile);
    int literal_num = tid % 9 + 1;
    auto any_left = [](const u8* t) { for (int i = 0; i < 34; i++) if (t[i] > 0) return true; return false; };

    if (literal_num <= 7 && tehai[tid + 1] > 0 && tehai[tid + 2] > 0) {
        u8 after[34]; memcpy(after, tehai, 34);
        after[tid] = 0; after[tid + 1] -= 1; after[tid + 2] -= 1;
        if (literal_num < 7) after[tid + 3] = 0;
        last_cans.can_chi_low = any_left(after);
    }
    if (literal_num >= 2 && literal_num <= 8 && tehai[tid - 1] > 0 && tehai[tid + 1] > 0) {
        u8 after[34]; memcpy(after, tehai, 34);
        after[tid] = 0; after[tid - 1] -= 1; after[tid + 1] -= 1;
        last_cans.can_chi_mid = any_left(after);
    }
    if (literal_num >= 3 && tehai[tid - 2] > 0 && tehai[tid - 1] > 0) {
        u8 after[34]; memcpy(after, tehai, 34);
        after[tid] = 0; after[tid - 2] -= 1; after[tid - 1] -= 1;
        if (literal_num > 3) after[tid - 3] = 0;
        last_cans.can_chi_high = any_left(after);
    }
}

// update.rs:875-878
void PlayerState::update_shanten() {
    shanten = std::max<i8>(shanten_all(tehai, tehai_len_div3), 0);
}

// update.rs:881-912
void PlayerState::update_shanten_discards() {
    ORC_ENSURE(last_cans.can_discard, "tehai is not 3n+2");
    memset(next_shanten_discards, 0, sizeof next_shanten_discards);
    memset(keep_shanten_discards, 0, sizeof keep_shanten_discards);
    has_next_shanten_discard = false;
    u8 th[34]; memcpy(th, tehai, 34);
    for (int tid = 0; tid < 34; tid++) {
        if (tehai[tid] == 0) continue;
        th[tid] -= 1;
        i8 after = shanten_all(th, tehai_len_div3);
        th[tid] += 1;
        if (after < shanten) { next_shanten_discards[tid] = true; has_next_shanten_discard = true; }
        else if (after == shanten) keep_shanten_discards[tid] = true;
    }
}

// update.rs:916-953
void PlayerState::update_waits_and_furiten() {
    ORC_ENSURE(!last_cans.can_discard, "tehai is not 3n+1");
    at_furiten = false;
    memset(waits, 0, sizeof waits);
    if (shanten > 0) return;
    for (int t = 0; t < 34; t++) {
        if (tehai[t] == 4) continue;
        u8 after[34]; memcpy(after, tehai, 34);
        after[t] += 1;
        if (shanten_all(after, tehai_len_div3) == -1) {
            if (discarded_tiles[t]) at_furiten = true;
            waits[t] = tiles_seen[t] < 4;
        }
    }
}

// update.rs:955-960
void PlayerState::update_doras_owned(int actor_rel, u8 tile) {
    doras_owned[actor_rel] += dora_factor[deaka(tile)];
    if (is_aka(tile)) doras_owned[actor_rel] += 1;
}

// update.rs:962-972
void PlayerState::update_rank() { rank = get_rank(scores); }
u8 PlayerState::get_rank(const i32* scores_rel) const {
    i32 abs_[4];
    for (int i = 0; i < 4; i++) abs_[(i + player_id) % 4] = scores_rel[i];  // rotate_right(player_id)
    u8 rbp[4];
    rankings(abs_, nullptr, rbp);
    return rbp[player_id];
}

// ---------------------------------------------------------------- action.rs
// chi_type.rs:10-25 — 0 low, 1 mid, 2 high
static int chi_type(const u8* consumed, u8 tile) {
    u8 a = deaka(consumed[0]), b = deaka(consumed[1]);
    u8 mn = std::min(a, b), mx = std::max(a, b), tid = deaka(tile);
    if (tid < mn) return 0;
    if (tid < mx) return 1;
    return 2;
}

// action.rs:213-227
void PlayerState::ensure_tiles_in_hand(const u8* tiles, int n) const {
    for (int i = 0; i < n; i++) {
        u8 t = tiles[i];
        ORC_ENSURE(t < T_UNK && tehai[deaka(t)] > 0, "tile is not in hand");
        if (is_aka(t)) ORC_ENSURE(akas_in_hand[t - T_5MR], "aka tile is not in hand");
    }
}

// action.rs:93-211
void PlayerState::validate_reaction(const Event& action) const {
    const ActionCandidate& cans = last_cans;
    if (action.type == EV_RYUKYOKU) { ORC_ENSURE(cans.can_ryukyoku, "cannot ryukyoku"); return; }
    if (action.type == EV_NONE) return;
    ORC_ENSURE(action.has_actor(), "action does not have actor and is not ryukyoku");
    ORC_ENSURE(action.actor == player_id, "actor is not self");

    switch (action.type) {
        case EV_DAHAI:
            ORC_ENSURE(cans.can_discard, "cannot discard");
            ensure_tiles_in_hand(&action.pai, 1);
            if (action.tsumogiri) {
                ORC_ENSURE(has_last_self_tsumo, "tsumogiri but the player has not dealt any tile yet");
                ORC_ENSURE(last_self_tsumo == action.pai, "cannot tsumogiri");
            }
            break;
        case EV_REACH: ORC_ENSURE(cans.can_riichi, "cannot riichi"); break;
        case EV_CHI: {
            ORC_ENSURE((action.target + 1) % 4 == action.actor, "chi from non-kamicha");
            ORC_ENSURE(has_last_kawa_tile && last_kawa_tile == action.pai, "chi target is not the last kawa tile");
            ensure_tiles_in_hand(action.consumed, 2);
            int ct = chi_type(action.consumed, action.pai);
            if (ct == 0) ORC_ENSURE(cans.can_chi_low, "cannot chi low");
            else if (ct == 1) ORC_ENSURE(cans.can_chi_mid, "cannot chi mid");
            else ORC_ENSURE(cans.can_chi_high, "cannot chi high");
            break;
        }
        case EV_PON:
            ORC_ENSURE(action.target != action.actor, "pon from itself");
            ORC_ENSURE(has_last_kawa_tile && last_kawa_tile == action.pai, "pon target is not the last kawa tile");
            ORC_ENSURE(cans.can_pon, "cannot pon");
            ensure_tiles_in_hand(action.consumed, 2);
            break;
        case EV_DAIMINKAN:
            ORC_ENSURE(action.target != action.actor, "daiminkan from itself");
            ORC_ENSURE(has_last_kawa_tile && last_kawa_tile == action.pai, "daiminkan target is not the last kawa tile");
            ORC_ENSURE(cans.can_daiminkan, "cannot daiminkan");
            ensure_tiles_in_hand(action.consumed, 3);
            break;
        case EV_KAKAN:
            ORC_ENSURE(cans.can_kakan, "cannot kakan");
            ORC_ENSURE(vec_contains(kakan_candidates, deaka(action.pai)), "cannot kakan this tile");
            ensure_tiles_in_hand(&action.pai, 1);
            break;
        case EV_ANKAN: {
            ORC_ENSURE(cans.can_ankan, "cannot ankan");
            u8 tile = deaka(action.consumed[0]);
            ORC_ENSURE(vec_contains(ankan_candidates, tile), "cannot ankan this tile");
            ensure_tiles_in_hand(action.consumed, 4);
            break;
        }
        case EV_HORA:
            if (action.target == player_id) ORC_ENSURE(cans.can_tsumo_agari, "cannot tsumo agari");
            else ORC_ENSURE(cans.can_ron_agari, "cannot ron agari");
            break;
        default: throw OrcError("unexpected action");
    }
}

// ---------------------------------------------------------------- agent_helper.rs
static void split_aka(const PlayerState& ps, bool* ret) {
    // agent_helper.rs:65-76 / 183-194
    if (ret[T_5M] && ps.akas_in_hand[0]) { ret[T_5MR] = true; ret[T_5M] = ps.tehai[T_5M] > 1; }
    if (ret[T_5P] && ps.akas_in_hand[1]) { ret[T_5PR] = true; ret[T_5P] = ps.tehai[T_5P] > 1; }
    if (ret[T_5S] && ps.akas_in_hand[2]) { ret[T_5SR] = true; ret[T_5S] = ps.tehai[T_5S] > 1; }
}

// agent_helper.rs:35-79
void PlayerState::discard_candidates_aka(bool* ret) const {
    ORC_ENSURE(last_cans.can_discard, "tehai is not 3n+2");
    for (int i = 0; i < 37; i++) ret[i] = false;
    if (riichi_accepted[0]) {
        ORC_ENSURE(has_last_self_tsumo, "riichi accepted without last self tsumo");
        ret[last_self_tsumo] = true;
        return;
    }
    for (int i = 0; i < 34; i++) {
        if (tehai[i] == 0) continue;
        if (riichi_declared[0]) ret[i] = shanten == 1 ? next_shanten_discards[i] : keep_shanten_discards[i];
        else ret[i] = !forbidden_tiles[i];
    }
    split_aka(*this, ret);
}

// agent_helper.rs:100-197
void PlayerState::discard_candidates_with_unconditional_tenpai_aka(bool* ret) const {
    ORC_ENSURE(last_cans.can_discard, "tehai is not 3n+2");
    for (int i = 0; i < 37; i++) ret[i] = false;
    if (tiles_left == 0 || shanten > 1 || (shanten == 1 && !has_next_shanten_discard)) return;

    if (has_last_self_tsumo) {
        if (waits[deaka(last_self_tsumo)]) return;
        if (riichi_accepted[0]) {
            if (!at_furiten) ret[last_self_tsumo] = true;
            return;
        }
    } else if (shanten_all(tehai, tehai_len_div3) == -1) {
        return;
    }

    const bool* tenpai_discards = shanten == 1 ? next_shanten_discards : keep_shanten_discards;
    for (int discard = 0; discard < 34; discard++) {
        if (!tenpai_discards[discard] || forbidden_tiles[discard]) continue;
        u8 t3n1[34]; memcpy(t3n1, tehai, 34);
        t3n1[discard] -= 1;
        for (int tsumo = 0; tsumo < 34; tsumo++) {
            u8 seen = tiles_seen[tsumo];
            if (tsumo == discard || t3n1[tsumo] == 4) continue;
            u8 t3n2[34]; memcpy(t3n2, t3n1, 34);
            t3n2[tsumo] += 1;
            if (shanten_all(t3n2, tehai_len_div3) > -1) continue;
            if (discarded_tiles[tsumo]) { ret[discard] = false; break; }
            if (seen == 4 || ret[discard]) continue;
            AgariCalc c = make_agari_calc(t3n2, (u8)tsumo, true);
            ret[discard] = c.has_yaku();
        }
    }
    split_aka(*this, ret);
}

// agent_helper.rs:88-96
void PlayerState::discard_candidates_with_unconditional_tenpai(bool* out34) const {
    bool full[37];
    discard_candidates_with_unconditional_tenpai_aka(full);
    for (int i = 0; i < 34; i++) out34[i] = full[i];
    out34[T_5M] |= full[T_5MR];
    out34[T_5S] |= full[T_5SR];
    out34[T_5P] |= full[T_5PR];
}

static const u8 YAOKYUU13[13] = {T_1M, T_9M, T_1P, T_9P, T_1S, T_9S, T_E, T_S, T_W, T_N, T_P, T_F, T_C};

// agent_helper.rs:201-206
u8 PlayerState::yaokyuu_kind_count() const {
    u8 n = 0;
    for (u8 t : YAOKYUU13) n += std::min<u8>(tehai[t], 1);
    return n;
}

// agent_helper.rs:262-271
bool PlayerState::rule_based_agari() const {
    if (!last_cans.can_agari()) return false;
    return rule_based_agari_slow(last_cans.can_ron_agari, rel(last_cans.target_actor));
}

// agent_helper.rs:273-368
bool PlayerState::rule_based_agari_slow(bool is_ron, int target_rel) const {
    if (!is_all_last || oya == 0 || rank < 3) return true;
    if (bakaze == T_W) {
        if (kyoku < 3) return true;
    } else {
        bool all_lt = true;
        for (int i = 0; i < 4; i++) if (scores[i] >= 30000) all_lt = false;
        if (all_lt) return true;
    }

    Point max_win_point;
    if (riichi_accepted[0]) {
        u8 full[34]; memcpy(full, tehai, 34);
        for (u8 t : ankan_overview[0]) full[t] += 4;
        // (tile, count) with count > 0, sort_unstable_by count descending.
        // NOTE: Rust's sort_unstable order among equal counts is unspecified; here ties
        // keep ascending tile id (what pattern-defeating quicksort yields for these tiny,
        // mostly-sorted inputs is insertion sort => stable). Only affects which ura
        // candidates are tried first among equally valuable ones.
        std::vector<std::pair<int, u8>> ordered;
        for (int t = 0; t < 34; t++) if (full[t] > 0) ordered.push_back({t, full[t]});
        std::stable_sort(ordered.begin(), ordered.end(), [](auto& l, auto& r) { return l.second > r.second; });
        u8 seen[34]; memcpy(seen, tiles_seen, 34);
        u8 ura[5]; int n_ura = 0;
        bool done = false;
        for (auto& pr : ordered) {
            if (done) break;
            u8 ura_ind = tile_prev((u8)pr.first);
            for (;;) {
                if (n_ura >= (int)dora_indicators.size()) { done = true; break; }
                if (seen[ura_ind] >= 4) break;
                ura[n_ura++] = ura_ind;
                seen[ura_ind] += 1;
            }
        }
        max_win_point = agari_points(is_ron, ura, n_ura);
    } else {
        max_win_point = agari_points(is_ron, nullptr, 0);
    }

    i32 exp_scores[4];
    for (int i = 0; i < 4; i++) exp_scores[i] = scores[i];
    if (is_ron) {
        exp_scores[0] += max_win_point.ron + kyotaku * 1000 + honba * 300;
        exp_scores[target_rel] -= max_win_point.ron + honba * 300;
    } else {
        exp_scores[0] += max_win_point.tsumo_total(false) + kyotaku * 1000 + honba * 300;
        for (int idx = 1; idx < 4; idx++) {
            if (idx == oya) exp_scores[idx] -= max_win_point.tsumo_oya + honba * 100;
            else exp_scores[idx] -= max_win_point.tsumo_ko + honba * 100;
        }
    }
    bool all_lt = true;
    for (int i = 0; i < 4; i++) if (exp_scores[i] >= 30000) all_lt = false;
    if (all_lt) return true;
    return get_rank(exp_scores) < 3;
}

// agent_helper.rs:377-462
Point PlayerState::agari_points(bool is_ron, const u8* ura, int n_ura) const {
    ORC_ENSURE((is_ron && last_cans.can_ron_agari) || last_cans.can_tsumo_agari, "cannot agari");
    if (!is_ron && can_w_riichi) return point_yakuman(oya == 0, 1);

    bool has_wt = is_ron ? has_last_kawa_tile : has_last_self_tsumo;
    ORC_ENSURE(has_wt, "cannot find the winning tile");
    u8 winning_tile = is_ron ? last_kawa_tile : last_self_tsumo;

    u8 additional_hans;
    if (is_ron) {
        additional_hans = (u8)riichi_accepted[0] + (u8)is_w_riichi + (u8)at_ippatsu + (u8)(tiles_left == 0) +
                          (u8)chankan_chance;
    } else {
        additional_hans = (u8)riichi_accepted[0] + (u8)is_w_riichi + (u8)at_ippatsu + (u8)is_menzen +
                          (u8)(tiles_left == 0 && !at_rinshan) + (u8)at_rinshan;
    }

    u8 th[34]; memcpy(th, tehai, 34);
    u8 final_doras = doras_owned[0];
    if (is_ron) {
        u8 tid = deaka(winning_tile);
        th[tid] += 1;
        final_doras += dora_factor[tid];
        if (is_aka(winning_tile)) final_doras += 1;
    }
    if (riichi_accepted[0]) {
        for (int i = 0; i < n_ura; i++) {
            u8 next = tile_next(ura[i]);
            u8 count = th[next];
            if (vec_contains(ankan_overview[0], next)) count += 4;
            final_doras += count;
        }
    }
    AgariCalc c = make_agari_calc(th, deaka(winning_tile), is_ron);
    Agari a = c.agari(additional_hans, final_doras);
    ORC_ENSURE(a.valid, "not a hora hand");
    return a.point(oya == 0);
}

// agent_helper.rs:467-503
i8 PlayerState::real_time_shanten() const {
    if (!last_cans.can_discard) return shanten;
    if (shanten > 0) return has_next_shanten_discard ? shanten - 1 : shanten;
    if (has_last_self_tsumo) return waits[deaka(last_self_tsumo)] ? -1 : 0;
    return shanten_all(tehai, tehai_len_div3);
}

}  // namespace orc
