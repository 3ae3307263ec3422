// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
#include "board.h"

#include <algorithm>

namespace orc {

// ================================================================ SHA3-256 (FIPS 202)
static inline u64 rotl64(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static void keccak_f1600(u64* s) {
    static const u64 RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        u64 C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(s[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= RC[round];
    }
}

void sha3_256(const u8* data, size_t len, u8* out32) {
    const size_t rate = 136;
    u64 s[25] = {0};
    u8 block[136];
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; i++) {
            u64 w = 0;
            for (int b = 0; b < 8; b++) w |= (u64)data[i * 8 + b] << (8 * b);
            s[i] ^= w;
        }
        keccak_f1600(s);
        data += rate;
        len -= rate;
    }
    memset(block, 0, rate);
    memcpy(block, data, len);
    block[len] ^= 0x06;
    block[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) {
        u64 w = 0;
        for (int b = 0; b < 8; b++) w |= (u64)block[i * 8 + b] << (8 * b);
        s[i] ^= w;
    }
    keccak_f1600(s);
    for (int i = 0; i < 32; i++) out32[i] = (u8)(s[i / 8] >> (8 * (i % 8)));
}

// ================================================================ ChaCha12 (rand_chacha 0.9.0)
// 64-bit block counter in words 12-13, stream id 0 in 14-15; output words consumed in order.
static inline u32 rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }

ChaCha12::ChaCha12(const u8* seed) {
    for (int i = 0; i < 8; i++)
        key[i] = seed[4 * i] | (seed[4 * i + 1] << 8) | (seed[4 * i + 2] << 16) | ((u32)seed[4 * i + 3] << 24);
}

u32 ChaCha12::next_u32() {
    if (pos >= 16) {
        u32 st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        for (int i = 0; i < 8; i++) st[4 + i] = key[i];
        st[12] = (u32)counter; st[13] = (u32)(counter >> 32); st[14] = 0; st[15] = 0;
        u32 x[16];
        memcpy(x, st, sizeof x);
#define QR(a, b, c, d)                                   \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);        \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);        \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);         \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
        for (int r = 0; r < 6; r++) {
            QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
            QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
        }
#undef QR
        for (int i = 0; i < 16; i++) buf[i] = x[i] + st[i];
        counter++;
        pos = 0;
    }
    return buf[pos++];
}

// ================================================================ shuffles
// rand 0.8: `for i in (1..len).rev() { swap(i, gen_range(0..=i)) }`, u32 widening-multiply
// with zone rejection. VERIFIED against log-viewer/index.example.html (tests/test_oracle_golden.py).
static u32 below_rand08(ChaCha12& rng, u32 range) {
    u32 zone = (range << __builtin_clz(range)) - 1;
    for (;;) {
        u32 v = rng.next_u32();
        u64 m = (u64)v * range;
        if ((u32)m <= zone) return (u32)(m >> 32);
    }
}
static void shuffle_rand08(u8* seq, int n, ChaCha12& rng) {
    for (int i = n - 1; i >= 1; i--) std::swap(seq[i], seq[below_rand08(rng, (u32)i + 1)]);
}

// rand 0.9.1 (Cargo.lock:1042): forward Fisher-Yates driven by IncreasingUniform, one
// Canon widening-multiply sample (single bias-reduction step) per chunk of indices.
// PARITY UNPINNED: no fixture in the reference pins this variant.
static u32 canon_below(ChaCha12& rng, u32 range) {
    u64 m = (u64)rng.next_u32() * range;
    u32 hi = (u32)(m >> 32), lo = (u32)m;
    if (lo > (u32)(0u - range)) {
        u32 new_hi = (u32)(((u64)rng.next_u32() * range) >> 32);
        u32 sum = lo + new_hi;
        if (sum < lo) hi += 1;
    }
    return hi;
}
static void shuffle_rand09(u8* seq, int len, ChaCha12& rng) {
    if (len <= 1) return;
    u32 n = 0, chunk = 0;
    int chunk_remaining = 1;  // IncreasingUniform::new(rng, 0)
    for (int i = 0; i < len; i++) {
        u32 next_n = n + 1;
        int next_cr;
        if (chunk_remaining == 0) {
            u32 product = next_n, current = next_n + 1;
            for (;;) {
                u64 p = (u64)product * current;
                if (p > 0xFFFFFFFFull) break;
                product = (u32)p;
                current += 1;
            }
            int remaining = (int)(current - next_n);
            chunk = canon_below(rng, product);
            next_cr = remaining - 1;
        } else {
            next_cr = chunk_remaining - 1;
        }
        u32 result;
        if (next_cr == 0) {
            result = chunk;
        } else {
            result = chunk % next_n;
            chunk /= next_n;
        }
        chunk_remaining = next_cr;
        n = next_n;
        std::swap(seq[i], seq[result]);
    }
}

// board.rs:99-110, 786-824
void make_wall(u64 nonce, u64 key, u8 kyoku, u8 honba, int shuffle_kind, u8* seq) {
    u8 msg[18];
    for (int i = 0; i < 8; i++) msg[i] = (u8)(nonce >> (8 * i));
    for (int i = 0; i < 8; i++) msg[8 + i] = (u8)(key >> (8 * i));
    msg[16] = kyoku;
    msg[17] = honba;
    u8 seed[32];
    sha3_256(msg, 18, seed);
    ChaCha12 rng(seed);
    for (int t = 0; t < 34; t++)
        for (int c = 0; c < 4; c++) seq[t * 4 + c] = (u8)t;
    seq[T_5M * 4] = T_5MR;
    seq[T_5P * 4] = T_5PR;
    seq[T_5S * 4] = T_5SR;
    if (shuffle_kind == SHUFFLE_RAND08) shuffle_rand08(seq, 136, rng);
    else shuffle_rand09(seq, 136, rng);
}

// board.rs:111-122
void Board::init_from_wall(const u8* seq) {
    for (int i = 0; i < 4; i++) memcpy(haipai[i], seq + 13 * i, 13);
    rinshan.assign(seq + 52, seq + 56);
    dora_indicators.assign(seq + 56, seq + 61);
    ura_indicators.assign(seq + 61, seq + 66);
    yama.assign(seq + 66, seq + 136);
}

// ================================================================ BoardState
BoardState::BoardState(const Board& b) : board(b) {
    oya = board.kyoku % 4;
    for (int i = 0; i < 4; i++) player_states[i] = PlayerState((u8)i);
    dora_indicators_full = board.dora_indicators;  // board.rs:127
}

int oracle_obs_rows(int version) {  // consts.rs:30-38
    if (version == 1) return 211;
    if (version >= 2 && version <= 4) return 217;
    throw OrcError("unsupported version");
}

// board.rs:680-782
void BoardState::encode_oracle_obs(u8 perspective, int version, float* out) const {
    const int rows = oracle_obs_rows(version);
    for (int i = 0; i < rows * 34; i++) out[i] = 0.f;
    auto assign = [&](int r, int c, float v) { out[r * 34 + c] = v; };
    auto fill = [&](int r, float v) { for (int c = 0; c < 34; c++) out[r * 34 + c] = v; };
    int idx = 0;
    for (int k = 1; k <= 3; k++) {  // .iter().cycle().skip(perspective + 1).take(3)
        const PlayerState& st = player_states[(perspective + k) % 4];
        for (int t = 0; t < 34; t++)
            for (int c = 0; c < st.tehai[t]; c++) assign(idx + c, t, 1.f);  // assign_rows
        idx += 4;
        for (int i = 0; i < 3; i++) if (st.akas_in_hand[i]) fill(idx + i, 1.f);
        idx += 3;
        const int n = st.shanten;
        if (version == 1) {
            for (int i = 0; i < n; i++) fill(idx + i, 1.f);  // fill_rows
            idx += 6;
        } else {
            fill(idx + n, 1.f);
            idx += 7;
            fill(idx, (float)n / 6.f);
            idx += 1;
        }
        for (int t = 0; t < 34; t++) if (st.waits[t]) assign(idx, t, 1.f);
        idx += 1;
        if (st.at_furiten) fill(idx, 1.f);
        idx += 1;
    }
    auto encode_tile = [&](int r, u8 tile) {
        assign(r, deaka(tile), 1.f);
        if (is_aka(tile)) fill(r + 1, 1.f);
    };
    {   // yama.iter().rev().take(tiles_left)
        int taken = 0;
        for (int i = (int)board.yama.size() - 1; i >= 0 && taken < tiles_left; i--, taken++) { encode_tile(idx, board.yama[i]); idx += 2; }
        idx += (69 - (int)tiles_left) * 2;
    }
    for (int i = (int)board.rinshan.size() - 1; i >= 0; i--) { encode_tile(idx, board.rinshan[i]); idx += 2; }
    idx += (4 - (int)board.rinshan.size()) * 2;
    for (int i = (int)dora_indicators_full.size() - 1; i >= 0; i--) { encode_tile(idx, dora_indicators_full[i]); idx += 2; }
    for (size_t i = 0; i < board.ura_indicators.size(); i++) { encode_tile(idx, board.ura_indicators[i]); idx += 2; }
    ORC_ENSURE(idx == rows, "encode_oracle_obs: row cursor mismatch");
}

// board.rs:141-161
Poll BoardState::poll(const Event in_reactions[4]) {
    Event reactions[4];
    for (int i = 0; i < 4; i++) reactions[i] = in_reactions[i];
    for (;;) {
        Poll p = step(reactions);
        if (p == POLL_INGAME) {
            for (auto& s : player_states)
                if (s.last_cans.can_act()) return p;
        } else {
            Event e; e.type = EV_END_KYOKU;
            log.push_back(e);
            for (int i = 0; i < 4; i++) board.scores[i] += kyoku_deltas[i];
            if (has_abortive_ryukyoku) can_renchan = true;
            return p;
        }
        for (int i = 0; i < 4; i++) reactions[i] = Event{};
    }
}

KyokuResult BoardState::end() const {
    KyokuResult r;
    r.kyoku = board.kyoku;
    r.can_renchan = can_renchan;
    r.has_hora = has_hora;
    r.has_abortive_ryukyoku = has_abortive_ryukyoku;
    r.kyotaku_left = board.kyotaku;
    for (int i = 0; i < 4; i++) r.scores[i] = board.scores[i];
    return r;
}

// board.rs:200-204
void BoardState::broadcast(const Event& ev) {
    for (auto& s : player_states) s.update(ev);
}

// board.rs:206-239
void BoardState::haipai() {
    Event sk;
    sk.type = EV_START_KYOKU;
    sk.bakaze = T_E + board.kyoku / 4;
    ORC_ENSURE(!board.dora_indicators.empty(), "insufficient dora indicators");
    sk.pai = board.dora_indicators.back();
    board.dora_indicators.pop_back();
    sk.kyoku = oya + 1;
    sk.honba = board.honba;
    sk.kyotaku = board.kyotaku;
    sk.oya = oya;
    for (int i = 0; i < 4; i++) sk.scores[i] = board.scores[i];
    memcpy(sk.tehais, board.haipai, sizeof sk.tehais);
    broadcast(sk);
    log.push_back(sk);

    ORC_ENSURE(!board.yama.empty(), "invalid yama: empty at init");
    u8 tile = board.yama.back();
    board.yama.pop_back();
    tiles_left -= 1;
    Event ts;
    ts.type = EV_TSUMO;
    ts.actor = oya;
    ts.pai = tile;
    broadcast(ts);
    log.push_back(ts);
}

// board.rs:241-294
void BoardState::exhaustive_ryukyoku() {
    i32 deltas[4] = {0, 0, 0, 0};
    can_renchan = player_states[oya].shanten == 0;

    bool has_nagashi = false;
    for (int i = 0; i < 4; i++) {
        if (!can_nagashi_mangan[i]) continue;
        has_nagashi = true;
        if (i == oya) {
            for (int j = 0; j < 4; j++) deltas[j] += (j == i) ? 12000 : -4000;
        } else {
            for (int j = 0; j < 4; j++) deltas[j] += (j == i) ? 8000 : (j == oya ? -4000 : -2000);
        }
    }
    if (!has_nagashi) {
        int tenpai[4], n = 0;
        for (int i = 0; i < 4; i++) if (player_states[i].shanten == 0) tenpai[n++] = i;
        i32 plus = 0, minus = 0;
        if (n == 1) { plus = 3000; minus = -1000; }
        else if (n == 2) { plus = 1500; minus = -1500; }
        else if (n == 3) { plus = 1000; minus = -3000; }
        if (plus > 0) {
            i32 dod[4] = {minus, minus, minus, minus};
            for (int k = 0; k < n; k++) dod[tenpai[k]] = plus;
            for (int j = 0; j < 4; j++) deltas[j] += dod[j];
        }
    }
    for (int j = 0; j < 4; j++) kyoku_deltas[j] += deltas[j];
    Event r;
    r.type = EV_RYUKYOKU;
    r.has_deltas = true;
    memcpy(r.deltas, deltas, sizeof deltas);
    log.push_back(r);
}

// board.rs:296-312
void BoardState::update_nagashi_mangan_and_four_wind(const Event& ev) {
    switch (ev.type) {
        case EV_DAHAI:
            if (!is_yaokyuu(ev.pai)) can_nagashi_mangan[ev.actor] = false;
            break;
        case EV_CHI: case EV_PON: case EV_DAIMINKAN:
            can_nagashi_mangan[ev.target] = false;
            can_four_wind = false;
            break;
        case EV_ANKAN: can_four_wind = false; break;
        default: break;
    }
}

// board.rs:314-340
bool BoardState::check_four_wind(u8 pai) {
    if (!(pai >= T_E && pai <= T_N)) {
        can_four_wind = false;
    } else if (player_states[tsumo_actor].can_w_riichi) {
        if (four_wind_tile >= 0) can_four_wind = four_wind_tile == pai;
        else four_wind_tile = pai;
    } else if (four_wind_tile >= 0) {
        if (four_wind_tile == pai) return true;
        can_four_wind = false;
    } else {
        throw OrcError("unexpected state when calculating four-wind");
    }
    return false;
}

// board.rs:342-351
void BoardState::check_riichi_accepted() {
    if (riichi_to_be_accepted < 0) return;
    u8 actor = (u8)riichi_to_be_accepted;
    riichi_to_be_accepted = -1;
    Event e;
    e.type = EV_REACH_ACCEPTED;
    e.actor = actor;
    broadcast(e);
    log.push_back(e);
    board.scores[actor] -= 1000;
    board.kyotaku += 1;
    accepted_riichis += 1;
}

// board.rs:353-364
void BoardState::add_new_dora() {
    ORC_ENSURE(!board.dora_indicators.empty(), "illegal kan: already 4 kans and this is the 5th");
    Event e;
    e.type = EV_DORA;
    e.pai = board.dora_indicators.back();
    board.dora_indicators.pop_back();
    broadcast(e);
    log.push_back(e);
}

// board.rs:366-471
void BoardState::handle_hora(u8 single_actor, u8 single_target, const Event reactions[4]) {
    has_hora = true;
    bool is_ron = single_actor != single_target;
    i32 honba_left = board.honba;
    i32 kyotaku_point = (i32)board.kyotaku * 1000;
    board.kyotaku = 0;

    int n_ura = 5 - (int)board.dora_indicators.size();
    const u8* ura = board.ura_indicators.data();

    bool has_point[4] = {false, false, false, false};
    Point points[4];
    for (int i = 0; i < 4; i++) {
        if (reactions[i].type != EV_HORA) continue;
        u8 actor = reactions[i].actor;
        can_renchan |= actor == oya;
        points[i] = player_states[actor].agari_points(is_ron, ura, n_ura);
        has_point[i] = true;
    }

    if (is_ron) {
        for (int k = 1; k <= 3; k++) {
            int actor = (single_target + k) % 4;
            if (!has_point[actor]) continue;
            const Point& point = points[actor];
            i32 deltas[4] = {0, 0, 0, 0};
            if (paos[actor] >= 0) {
                int pao_target = paos[actor];
                deltas[pao_target] = -point.ron / 2 - honba_left * 300;
                deltas[single_target] -= point.ron / 2;
            } else {
                deltas[single_target] = -point.ron - honba_left * 300;
            }
            deltas[actor] = point.ron + kyotaku_point + honba_left * 300;
            kyotaku_point = 0;
            honba_left = 0;
            for (int j = 0; j < 4; j++) kyoku_deltas[j] += deltas[j];
            Event h;
            h.type = EV_HORA;
            h.actor = (u8)actor;
            h.target = single_target;
            h.has_deltas = true;
            memcpy(h.deltas, deltas, sizeof deltas);
            if (player_states[actor].riichi_accepted[0]) {
                h.n_ura = n_ura;
                for (int j = 0; j < n_ura; j++) h.ura_markers[j] = ura[j];
            }
            log.push_back(h);
        }
        return;
    }

    ORC_ENSURE(has_point[single_actor], "tsumo hora without point");
    const Point& point = points[single_actor];
    i32 deltas[4] = {0, 0, 0, 0};
    if (paos[single_actor] >= 0) {
        deltas[paos[single_actor]] = -point.ron - honba_left * 300;
    } else {
        for (int j = 0; j < 4; j++) deltas[j] = -point.tsumo_ko - honba_left * 100;
        if (single_actor != oya) deltas[oya] = -point.tsumo_oya - honba_left * 100;
    }
    deltas[single_actor] = point.tsumo_total(single_actor == oya) + kyotaku_point + honba_left * 300;
    for (int j = 0; j < 4; j++) kyoku_deltas[j] += deltas[j];
    Event h;
    h.type = EV_HORA;
    h.actor = single_actor;
    h.target = single_target;
    h.has_deltas = true;
    memcpy(h.deltas, deltas, sizeof deltas);
    if (player_states[single_actor].riichi_accepted[0]) {
        h.n_ura = n_ura;
        for (int j = 0; j < n_ura; j++) h.ura_markers[j] = ura[j];
    }
    log.push_back(h);
}

// board.rs:473-499
void BoardState::update_paos(const Event& ev) {
    if (!(ev.type == EV_PON || ev.type == EV_DAIMINKAN) || !is_jihai(ev.pai)) return;
    u8 jihais = 0;
    const PlayerState& ps = player_states[ev.actor];
    for (u8 t : ps.pons) if (t >= T_E) jihais |= 1 << (t - T_E);
    for (u8 t : ps.minkans) if (t >= T_E) jihais |= 1 << (t - T_E);
    bool daisangen = (jihais & 0b1110000) == 0b1110000;
    bool daisuushi = (jihais & 0b0001111) == 0b0001111;
    bool is_sangen = ev.pai >= T_P && ev.pai <= T_C;
    bool is_wind = ev.pai >= T_E && ev.pai <= T_N;
    if ((daisangen && is_sangen) || (daisuushi && is_wind)) paos[ev.actor] = ev.target;
}

// board.rs:502-509
void BoardState::abortive_ryukyoku() {
    Event r;
    r.type = EV_RYUKYOKU;
    r.has_deltas = true;
    log.push_back(r);
    has_abortive_ryukyoku = true;
}

// board.rs:511-678
Poll BoardState::step(const Event reactions[4]) {
    if (tiles_left == 70) {
        haipai();
        return POLL_INGAME;
    }
    if (accepted_riichis == 4) {
        abortive_ryukyoku();
        return POLL_END;
    }
    for (int a = 0; a < 4; a++) player_states[a].validate_reaction(reactions[a]);

    auto prio = [](const Event& e) {
        switch (e.type) {
            case EV_HORA: return 0;
            case EV_DAIMINKAN: case EV_PON: return 1;
            case EV_NONE: return 3;
            default: return 2;
        }
    };
    int best = 0;
    for (int a = 1; a < 4; a++) if (prio(reactions[a]) < prio(reactions[best])) best = a;
    const Event& ev = reactions[best];

    if (check_four_kan && ev.type != EV_HORA) {
        abortive_ryukyoku();
        return POLL_END;
    }

    update_nagashi_mangan_and_four_wind(ev);

    switch (ev.type) {
        case EV_NONE: {
            if (tiles_left == 0) {
                exhaustive_ryukyoku();
                return POLL_END;
            }
            check_riichi_accepted();
            u8 tile;
            if (deal_from_rinshan) {
                deal_from_rinshan = false;
                ORC_ENSURE(!board.rinshan.empty(), "illegal kan: already 4 kans and this is the 5th");
                tile = board.rinshan.back();
                board.rinshan.pop_back();
            } else {
                ORC_ENSURE(!board.yama.empty(), "tiles left > 0 but yama is empty");
                tile = board.yama.back();
                board.yama.pop_back();
            }
            tiles_left -= 1;
            Event ts;
            ts.type = EV_TSUMO;
            ts.actor = tsumo_actor;
            ts.pai = tile;
            if (need_new_dora_at_tsumo) {
                need_new_dora_at_tsumo = false;
                add_new_dora();
            }
            broadcast(ts);
            log.push_back(ts);
            break;
        }
        case EV_DAHAI: {
            if (need_new_dora_at_discard) {
                need_new_dora_at_discard = false;
                add_new_dora();
            }
            broadcast(ev);
            log.push_back(ev);
            tsumo_actor = (ev.actor + 1) % 4;
            if (can_four_wind && check_four_wind(ev.pai)) {
                abortive_ryukyoku();
                return POLL_END;
            }
            if (kans == 4) {
                bool all_lt4 = true;
                for (auto& s : player_states) if (s.kans_count() >= 4) all_lt4 = false;
                if (all_lt4) check_four_kan = true;
            }
            break;
        }
        case EV_CHI: case EV_PON:
            check_riichi_accepted();
            broadcast(ev);
            log.push_back(ev);
            break;
        case EV_ANKAN:
            if (need_new_dora_at_discard) {
                need_new_dora_at_discard = false;
                add_new_dora();
            }
            broadcast(ev);
            log.push_back(ev);
            add_new_dora();
            tsumo_actor = ev.actor;
            deal_from_rinshan = true;
            kans += 1;
            break;
        case EV_DAIMINKAN: case EV_KAKAN:
            if (need_new_dora_at_discard) need_new_dora_at_tsumo = true;
            check_riichi_accepted();
            broadcast(ev);
            log.push_back(ev);
            need_new_dora_at_discard = true;
            tsumo_actor = ev.actor;
            deal_from_rinshan = true;
            kans += 1;
            break;
        case EV_REACH:
            broadcast(ev);
            log.push_back(ev);
            riichi_to_be_accepted = ev.actor;
            break;
        case EV_HORA:
            handle_hora(ev.actor, ev.target, reactions);
            return POLL_END;
        case EV_RYUKYOKU:
            abortive_ryukyoku();
            return POLL_END;
        default:
            throw OrcError("unexpected event");
    }
    update_paos(ev);
    return POLL_INGAME;
}

// ================================================================ agent (mortal.rs)
// obs_repr.rs mask writes: 423-427, 445-447, 480-559
void legal_mask(const PlayerState& st, bool at_kan_select, u8* mask) {
    memset(mask, 0, 46);
    const ActionCandidate& cans = st.last_cans;
    if (cans.can_pass()) {
        ORC_ENSURE(st.has_last_kawa_tile, "building chi/pon/daiminkan/ron feature without any kawa tile");
        if (!at_kan_select) mask[45] = 1;
        else if (cans.can_daiminkan) mask[deaka(st.last_kawa_tile)] = 1;
    }
    if (cans.can_discard) {
        bool dc[37];
        st.discard_candidates_aka(dc);
        if (!at_kan_select)
            for (int t = 0; t < 37; t++) if (dc[t]) mask[t] = 1;
    }
    if (cans.can_riichi && !at_kan_select) mask[37] = 1;
    if (cans.can_chi_low && !at_kan_select) mask[38] = 1;
    if (cans.can_chi_mid && !at_kan_select) mask[39] = 1;
    if (cans.can_chi_high && !at_kan_select) mask[40] = 1;
    if (cans.can_pon && !at_kan_select) mask[41] = 1;
    if (cans.can_daiminkan && !at_kan_select) mask[42] = 1;
    if (cans.can_ankan) {
        if (at_kan_select) for (u8 t : st.ankan_candidates) mask[t] = 1;
        else mask[42] = 1;
    }
    if (cans.can_kakan) {
        if (at_kan_select) for (u8 t : st.kakan_candidates) mask[t] = 1;
        else mask[42] = 1;
    }
    if (cans.can_agari() && !at_kan_select) mask[43] = 1;
    if (cans.can_ryukyoku && !at_kan_select) mask[44] = 1;
}

static inline bool contains(const std::vector<u8>& v, u8 x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// mortal.rs:338-573
Event decode_action(const PlayerState& st, u8 actor, int action, int kan_select_action) {
    const ActionCandidate& cans = st.last_cans;
    const bool* akas = st.akas_in_hand;
    Event ev;
    auto aka_for = [&](u8 pai, u8 a_m, u8 b_m) {
        // the match arms list the same two ranks for each of the three suits
        for (int k = 0; k < 3; k++)
            if (pai == a_m + 9 * k || pai == b_m + 9 * k) return akas[k];
        return false;
    };
    if (action >= 0 && action <= 36) {
        ORC_ENSURE(cans.can_discard, "failed discard check");
        ev.type = EV_DAHAI;
        ev.actor = actor;
        ev.pai = (u8)action;
        ev.tsumogiri = st.has_last_self_tsumo && st.last_self_tsumo == ev.pai;
    } else if (action == 37) {
        ORC_ENSURE(cans.can_riichi, "failed riichi check");
        ev.type = EV_REACH;
        ev.actor = actor;
    } else if (action == 38) {
        ORC_ENSURE(cans.can_chi_low, "failed chi low check");
        ORC_ENSURE(st.has_last_kawa_tile, "invalid state: no last kawa tile");
        u8 pai = st.last_kawa_tile;
        u8 first = tile_next(pai);
        bool ak = aka_for(pai, 2, 3);  // 3m|4m
        ev.type = EV_CHI; ev.actor = actor; ev.target = cans.target_actor; ev.pai = pai;
        ev.consumed[0] = ak ? akaize(first) : first;
        ev.consumed[1] = ak ? akaize(tile_next(first)) : tile_next(first);
    } else if (action == 39) {
        ORC_ENSURE(cans.can_chi_mid, "failed chi mid check");
        ORC_ENSURE(st.has_last_kawa_tile, "invalid state: no last kawa tile");
        u8 pai = st.last_kawa_tile;
        bool ak = aka_for(pai, 3, 5);  // 4m|6m
        ev.type = EV_CHI; ev.actor = actor; ev.target = cans.target_actor; ev.pai = pai;
        ev.consumed[0] = ak ? akaize(tile_prev(pai)) : tile_prev(pai);
        ev.consumed[1] = ak ? akaize(tile_next(pai)) : tile_next(pai);
    } else if (action == 40) {
        ORC_ENSURE(cans.can_chi_high, "failed chi high check");
        ORC_ENSURE(st.has_last_kawa_tile, "invalid state: no last kawa tile");
        u8 pai = st.last_kawa_tile;
        u8 last = tile_prev(pai);
        bool ak = aka_for(pai, 5, 6);  // 6m|7m
        ev.type = EV_CHI; ev.actor = actor; ev.target = cans.target_actor; ev.pai = pai;
        ev.consumed[0] = ak ? akaize(tile_prev(last)) : tile_prev(last);
        ev.consumed[1] = ak ? akaize(last) : last;
    } else if (action == 41) {
        ORC_ENSURE(cans.can_pon, "failed pon check");
        ORC_ENSURE(st.has_last_kawa_tile, "invalid state: no last kawa tile");
        u8 pai = st.last_kawa_tile;
        bool ak = aka_for(pai, 4, 4);  // 5m (non-aka only)
        ev.type = EV_PON; ev.actor = actor; ev.target = cans.target_actor; ev.pai = pai;
        ev.consumed[0] = ak ? akaize(pai) : deaka(pai);
        ev.consumed[1] = deaka(pai);
    } else if (action == 42) {
        ORC_ENSURE(cans.can_daiminkan || cans.can_ankan || cans.can_kakan, "failed kan check");
        u8 tile;
        if (kan_select_action >= 0) {
            tile = (u8)kan_select_action;
            ORC_ENSURE(contains(st.ankan_candidates, tile) || contains(st.kakan_candidates, tile),
                       "kan choice not in kan candidates");
        } else if (cans.can_daiminkan) {
            ORC_ENSURE(st.has_last_kawa_tile, "invalid state: no last kawa tile");
            tile = st.last_kawa_tile;
        } else if (cans.can_ankan) {
            tile = st.ankan_candidates[0];
        } else {
            tile = st.kakan_candidates[0];
        }
        if (cans.can_daiminkan) {
            ev.type = EV_DAIMINKAN; ev.actor = actor; ev.target = cans.target_actor; ev.pai = tile;
            if (is_aka(tile)) { ev.consumed[0] = ev.consumed[1] = ev.consumed[2] = deaka(tile); }
            else { ev.consumed[0] = akaize(tile); ev.consumed[1] = tile; ev.consumed[2] = tile; }
        } else if (cans.can_ankan && contains(st.ankan_candidates, deaka(tile))) {
            ev.type = EV_ANKAN; ev.actor = actor;
            ev.consumed[0] = akaize(tile); ev.consumed[1] = ev.consumed[2] = ev.consumed[3] = tile;
        } else {
            bool ak = aka_for(tile, 4, 4);
            ev.type = EV_KAKAN; ev.actor = actor;
            if (ak) { ev.pai = akaize(tile); ev.consumed[0] = ev.consumed[1] = ev.consumed[2] = deaka(tile); }
            else { ev.pai = deaka(tile); ev.consumed[0] = akaize(tile); ev.consumed[1] = ev.consumed[2] = deaka(tile); }
        }
    } else if (action == 43) {
        ORC_ENSURE(cans.can_agari(), "failed hora check");
        ev.type = EV_HORA; ev.actor = actor; ev.target = cans.target_actor;
    } else if (action == 44) {
        ORC_ENSURE(cans.can_ryukyoku, "failed ryukyoku check");
        ev.type = EV_RYUKYOKU;
    } else {
        ev.type = EV_NONE;
    }
    return ev;
}

// ================================================================ Game (game.rs)
void Game::poll() {
    if (ended) return;
    if (!kyoku_started) {
        bool any30k = false;
        for (int i = 0; i < 4; i++) if (scores[i] >= 30000) any30k = true;
        if (kyoku >= length + 4 || (kyoku >= length && !in_renchan && any30k)) {
            ended = true;
            return;
        }
        Board nb;
        nb.kyoku = kyoku; nb.honba = honba; nb.kyotaku = kyotaku;
        for (int i = 0; i < 4; i++) nb.scores[i] = scores[i];
        u8 seq[136];
        if (!(wall_override && wall_override(kyoku, honba, seq)))
            make_wall(seed_nonce, seed_key, kyoku, honba, shuffle_kind, seq);
        nb.init_from_wall(seq);
        delete board;
        board = new BoardState(nb);
        kyoku_started = true;
    }

    Event reactions[4];
    for (int i = 0; i < 4; i++) { reactions[i] = last_reactions[i]; last_reactions[i] = Event{}; }
    Poll p = board->poll(reactions);
    if (p == POLL_INGAME) return;

    kyoku_started = false;
    in_renchan = false;
    KyokuResult kr = board->end();
    kyotaku = kr.kyotaku_left;
    for (int i = 0; i < 4; i++) scores[i] = kr.scores[i];
    game_log.push_back(board->log);

    for (int i = 0; i < 4; i++) if (scores[i] < 0) { ended = true; return; }

    if (kr.has_abortive_ryukyoku) { honba += 1; return poll(); }
    if (!kr.can_renchan) {
        kyoku += 1;
        if (kr.has_hora) honba = 0; else honba += 1;
        return poll();
    }
    int oya = kr.kyoku % 4;
    if (kr.kyoku >= length - 1 && scores[oya] >= 30000) {
        int top = 0;
        for (int i = 1; i < 4; i++) if (-kr.scores[i] < -kr.scores[top]) top = i;  // first minimum of -s
        if (top == oya) { ended = true; return; }
    }
    in_renchan = true;
    honba += 1;
    return poll();
}

bool Game::commit(const AgentConfig cfgs[4], const PolicyFn pols[4], std::vector<int>* trace) {
    if (ended) {
        if (kyotaku > 0) {
            int top = 0;
            for (int i = 1; i < 4; i++) if (-scores[i] < -scores[top]) top = i;
            scores[top] += (i32)kyotaku * 1000;
            // NOTE game.rs:181-184 does not clear self.kyotaku; commit() is called once.
        }
        return true;
    }
    for (u8 seat = 0; seat < 4; seat++) {
        const PlayerState& st = board->player_states[seat];
        const ActionCandidate& cans = st.last_cans;
        if (!cans.can_act()) continue;
        const AgentConfig& cfg = cfgs[seat];

        // mortal.rs:210-242 quick eval
        if (cfg.enable_quick_eval && cans.can_discard && !cans.can_riichi && !cans.can_tsumo_agari &&
            !cans.can_ankan && !cans.can_kakan && !cans.can_ryukyoku) {
            bool dc[37];
            st.discard_candidates_aka(dc);
            int only = -1, cnt = 0;
            for (int t = 0; t < 37; t++) if (dc[t]) { if (cnt == 0) only = t; cnt++; }
            if (cnt == 1) {
                Event ev;
                ev.type = EV_DAHAI; ev.actor = seat; ev.pai = (u8)only;
                ev.tsumogiri = st.has_last_self_tsumo && st.last_self_tsumo == ev.pai;
                last_reactions[seat] = ev;
                if (trace) { trace->push_back(seat); trace->push_back(only); trace->push_back(-2); }
                continue;
            }
        }
        // mortal.rs:244-250
        bool need_kan_select;
        if (!cans.can_ankan && !cans.can_kakan) need_kan_select = false;
        else if (!cfg.enable_quick_eval) need_kan_select = true;
        else need_kan_select = st.ankan_candidates.size() + st.kakan_candidates.size() > 1;

        Scene sc;
        sc.table = table; sc.seat = seat; sc.step_idx = step_idx; sc.state = &st; sc.board = board;
        u8 mask[46];
        int kan_action = -1;
        if (need_kan_select) {
            sc.is_kan_select = true;
            legal_mask(st, true, mask);
            kan_action = pols[seat](sc, mask, nullptr);
            ORC_ENSURE(kan_action >= 0 && kan_action < 46 && mask[kan_action], "policy returned illegal kan-select action");
        }
        sc.is_kan_select = false;
        legal_mask(st, false, mask);
        float q[46];
        for (int i = 0; i < 46; i++) q[i] = 0.f;
        int action = pols[seat](sc, mask, q);
        ORC_ENSURE(action >= 0 && action < 46 && mask[action], "policy returned illegal action");
        // mortal.rs:319-336
        if (cfg.enable_rule_based_agari_guard && action == 43 && !st.rule_based_agari()) {
            int best = -1;
            for (int i = 0; i < 46; i++) {
                if (i == 43 || !mask[i]) continue;
                if (best < 0 || q[i] >= q[best]) best = i;  // max_by returns the last maximum
            }
            ORC_ENSURE(best >= 0, "agari guard with no alternative");
            action = best;
        }
        if (trace) { trace->push_back(seat); trace->push_back(action); trace->push_back(kan_action); }
        last_reactions[seat] = decode_action(st, seat, action, action == 42 ? kan_action : -1);
    }
    step_idx++;
    return false;
}

// ================================================================ test policies
u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
u64 policy_hash(u64 nonce, u64 key, u64 table, u64 step_idx, u32 seat, u32 kan) {
    u64 h = splitmix64(nonce);
    h = splitmix64(h ^ key);
    h = splitmix64(h ^ table);
    h = splitmix64(h ^ step_idx);
    h = splitmix64(h ^ (u64)(seat * 2 + kan));
    return h;
}

static int kth_set(const u8* mask, int lo, int hi, int k) {
    for (int i = lo; i < hi; i++)
        if (mask[i]) { if (k == 0) return i; k--; }
    return -1;
}
static int count_set(const u8* mask, int lo, int hi) {
    int n = 0;
    for (int i = lo; i < hi; i++) n += mask[i] ? 1 : 0;
    return n;
}

// kind 0: uniform over the legal mask.
// kind 1: "greedy": agari whenever legal; riichi with p=3/4; if any non-discard option is
//   legal (calls / kan / ryukyoku / pass) pick uniformly among {those} with p=1/2 when discards
//   are also legal; discards prefer next-shanten, then keep-shanten, then any legal discard,
//   uniformly inside the preferred class. Defined on (mask, keep/next planes) only so the
//   CUDA test policy (mortal_b200/csrc/policy_test.cu) computes the identical choice.
int test_policy(int kind, const Scene& sc, u64 nonce, u64 key, const u8* mask) {
    u64 h = policy_hash(nonce, key, (u64)sc.table, sc.step_idx, sc.seat, sc.is_kan_select ? 1 : 0);
    if (kind == 2) {
        // kind 2 = kind 1 with the hash taken from the legal mask alone: a function of what an engine is handed
        // (mask + the keep / next-shanten planes of the observation), so the CPU arm, the device test policy and a
        // react_batch engine behind the plugin API all play the very same games (bench.py)
        u64 bits = 0;
        for (int i = 0; i < 46; i++) if (mask[i]) bits |= 1ull << i;
        h = splitmix64(bits);
        kind = 1;
    }
    if (kind == 0 || sc.is_kan_select) {
        int n = count_set(mask, 0, 46);
        return kth_set(mask, 0, 46, (int)(h % (u64)n));
    }
    if (mask[43]) return 43;
    u64 h2 = splitmix64(h);
    if (mask[37] && (h2 & 3) != 0) return 37;
    int n_disc = count_set(mask, 0, 37);
    u8 other[46];
    memcpy(other, mask, 46);
    for (int i = 0; i < 37; i++) other[i] = 0;
    other[37] = 0;
    int n_other = count_set(other, 0, 46);
    u64 h3 = splitmix64(h2);
    if (n_other > 0 && (n_disc == 0 || (h3 & 1))) return kth_set(other, 0, 46, (int)((h3 >> 1) % (u64)n_other));
    // discards
    const PlayerState& st = *sc.state;
    u8 pref[37];
    for (int pass = 0; pass < 2; pass++) {
        const bool* plane = pass == 0 ? st.next_shanten_discards : st.keep_shanten_discards;
        int n = 0;
        for (int t = 0; t < 37; t++) {
            pref[t] = mask[t] && plane[deaka((u8)t)];
            n += pref[t];
        }
        if (n > 0) return kth_set(pref, 0, 37, (int)((h3 >> 1) % (u64)n));
    }
    return kth_set(mask, 0, 37, (int)((h3 >> 1) % (u64)n_disc));
}

}  // namespace orc
