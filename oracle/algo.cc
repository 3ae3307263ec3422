// ORACLE — TEST INFRASTRUCTURE ONLY (see mj.h).
// Restates libriichi algo/{shanten,agari,point}.rs + rankings.rs + tile.rs priorities.
#include "mj.h"

#include <algorithm>
#include <cstdio>
#include <unordered_map>

namespace orc {

// ---------------------------------------------------------------- tile.rs:20-27
static const u8 DISCARD_PRIORITIES[38] = {
    6, 5, 4, 3, 2, 3, 4, 5, 6,  // m
    6, 5, 4, 3, 2, 3, 4, 5, 6,  // p
    6, 5, 4, 3, 2, 3, 4, 5, 6,  // s
    7, 7, 7, 7, 7, 7, 7,        // z
    1, 1, 1,                    // aka
    0,                          // unknown
};

// tile.rs:177-185: compare priority, ties broken by REVERSED id order.
int cmp_discard_priority(u8 l, u8 r) {
    int pl = DISCARD_PRIORITIES[l], pr = DISCARD_PRIORITIES[r];
    if (pl != pr) return pl < pr ? -1 : 1;
    if (r != l) return r < l ? -1 : 1;
    return 0;
}

// ---------------------------------------------------------------- tables
static std::vector<std::array<u8, 10>> g_suhai, g_jihai;
static std::unordered_map<u32, std::vector<u32>> g_agari;
static bool g_ready = false;

static std::vector<u8> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw OrcError("oracle: cannot open table file " + path);
    std::vector<u8> buf;
    u8 tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return buf;
}

// shanten.rs:27-44 — 5 bytes per 10-nibble row, low nibble first.
static std::vector<std::array<u8, 10>> unpack_rows(const std::vector<u8>& raw, size_t length) {
    std::vector<std::array<u8, 10>> ret;
    ret.reserve(length);
    std::array<u8, 10> entry{};
    for (size_t i = 0; i < raw.size(); i++) {
        entry[i * 2 % 10] = raw[i] & 0xF;
        entry[i * 2 % 10 + 1] = (raw[i] >> 4) & 0xF;
        if ((i + 1) % 5 == 0) ret.push_back(entry);
    }
    if (ret.size() != length) throw OrcError("oracle: shanten table has wrong length");
    return ret;
}

void tables_init(const char* data_dir) {
    if (g_ready) return;
    std::string d(data_dir);
    g_jihai = unpack_rows(read_file(d + "/shanten_jihai.bin"), 78032);     // shanten.rs:11
    g_suhai = unpack_rows(read_file(d + "/shanten_suhai.bin"), 1940777);   // shanten.rs:12
    // agari.rs:24-37: repeated (u32 key LE, u8 n, n x u32 LE)
    std::vector<u8> raw = read_file(d + "/agari.bin");
    size_t p = 0;
    auto rd32 = [&](void) {
        u32 v = raw[p] | (raw[p + 1] << 8) | (raw[p + 2] << 16) | ((u32)raw[p + 3] << 24);
        p += 4;
        return v;
    };
    for (int i = 0; i < 9362; i++) {  // agari.rs:22
        u32 key = rd32();
        int n = raw[p++];
        std::vector<u32> v;
        for (int j = 0; j < n; j++) v.push_back(rd32());
        if (g_agari.count(key)) throw OrcError("oracle: duplicated agari key");
        g_agari[key] = v;
    }
    if (p != raw.size()) throw OrcError("oracle: trailing bytes in agari table");
    g_ready = true;
}
bool tables_ready() { return g_ready; }

int agari_table_lookup(u32 key, u32* divs4) {
    auto it = g_agari.find(key);
    if (it == g_agari.end()) return -1;
    int n = (int)it->second.size();
    for (int i = 0; i < n && i < 4; i++) divs4[i] = it->second[i];
    return n;
}

// ---------------------------------------------------------------- shanten.rs
typedef std::array<u8, 10> Row;
static inline Row row_or_default(const std::vector<Row>& tab, size_t idx) {
    if (idx < tab.size()) return tab[idx];
    return Row{};
}

// shanten.rs:51-69
static void add_suhai(Row& lhs, size_t index, size_t m) {
    Row tab = row_or_default(g_suhai, index);
    for (size_t j = 5 + m; j >= 5; j--) {
        u8 sht = std::min<u8>(lhs[j] + tab[0], lhs[0] + tab[j]);
        for (size_t k = 5; k < j; k++) {
            sht = std::min<u8>(sht, lhs[k] + tab[j - k]);
            sht = std::min<u8>(sht, lhs[j - k] + tab[k]);
        }
        lhs[j] = sht;
    }
    for (size_t jj = m + 1; jj-- > 0;) {
        size_t j = jj;
        u8 sht = lhs[j] + tab[0];
        for (size_t k = 0; k < j; k++) sht = std::min<u8>(sht, lhs[k] + tab[j - k]);
        lhs[j] = sht;
    }
}

// shanten.rs:71-80
static void add_jihai(Row& lhs, size_t index, size_t m) {
    Row tab = row_or_default(g_jihai, index);
    size_t j = m + 5;
    u8 sht = std::min<u8>(lhs[j] + tab[0], lhs[0] + tab[j]);
    for (size_t k = 5; k < j; k++) {
        sht = std::min<u8>(sht, lhs[k] + tab[j - k]);
        sht = std::min<u8>(sht, lhs[j - k] + tab[k]);
    }
    lhs[j] = sht;
}

// shanten.rs:82-84
static size_t sum_tiles(const u8* t, int n) {
    size_t acc = 0;
    for (int i = 0; i < n; i++) acc = acc * 5 + t[i];
    return acc;
}

// shanten.rs:88-100
i8 shanten_normal(const u8* tiles, u8 len_div3) {
    size_t m = len_div3;
    Row ret = row_or_default(g_suhai, sum_tiles(tiles, 9));
    add_suhai(ret, sum_tiles(tiles + 9, 9), m);
    add_suhai(ret, sum_tiles(tiles + 18, 9), m);
    add_jihai(ret, sum_tiles(tiles + 27, 7), m);
    return (i8)ret[5 + m] - 1;
}

// shanten.rs:103-115
i8 shanten_chitoi(const u8* tiles) {
    int pairs = 0, kinds = 0;
    for (int i = 0; i < 34; i++)
        if (tiles[i] > 0) {
            kinds++;
            if (tiles[i] >= 2) pairs++;
        }
    int redunct = kinds >= 7 ? 0 : 7 - kinds;
    return (i8)(7 - pairs + redunct - 1);
}

static const u8 YAOKYUU[13] = {T_1M, T_9M, T_1P, T_9P, T_1S, T_9S, T_E, T_S, T_W, T_N, T_P, T_F, T_C};

// shanten.rs:118-135
i8 shanten_kokushi(const u8* tiles) {
    int pairs = 0, kinds = 0;
    for (u8 t : YAOKYUU)
        if (tiles[t] > 0) {
            kinds++;
            if (tiles[t] >= 2) pairs++;
        }
    int redunct = pairs > 0 ? 1 : 0;
    return (i8)(14 - kinds - redunct - 1);
}

// shanten.rs:138-150
i8 shanten_all(const u8* tiles, u8 len_div3) {
    i8 s = shanten_normal(tiles, len_div3);
    if (s <= 0 || len_div3 < 4) return s;
    s = std::min(s, shanten_chitoi(tiles));
    if (s > 0) return std::min(s, shanten_kokushi(tiles));
    return s;
}

// ---------------------------------------------------------------- point.rs
// point.rs:13-84. The reference is an explicit match table; point.rs:120-154 (its
// own test) proves it equals this closed form on every (fu, han) it accepts.
// Combinations the reference panics on raise OrcError here.
Point point_calc(bool is_oya, u8 fu, u8 han) {
    auto bad = [&]() {
        throw OrcError("impossible combination of " + std::to_string(fu) + " fu and " + std::to_string(han) + " han");
    };
    i32 base;
    if (han >= 13) base = 8000;
    else if (han >= 11) base = 6000;
    else if (han >= 8) base = 4000;
    else if (han >= 6) base = 3000;
    else if (han == 5) base = 2000;
    else {
        // (_, 5) | (40.., 4) | (70.., 3) are mangan; everything else must be in the explicit table
        if (han == 0) bad();
        bool mangan = (han == 4 && fu >= 40) || (han == 3 && fu >= 70);
        if (mangan) base = 2000;
        else {
            bool listed = false;
            switch (fu) {
                case 20: listed = han >= 2 && han <= 4; break;
                case 25: listed = han >= 2 && han <= 4; break;
                case 30: listed = han >= 1 && han <= 4; break;
                case 40: listed = han >= 1 && han <= 3; break;
                case 50: listed = han >= 1 && han <= 3; break;
                case 60: listed = han >= 1 && han <= 3; break;
                case 70: listed = han >= 1 && han <= 2; break;
                case 80: listed = han >= 1 && han <= 2; break;
                case 90: listed = han >= 1 && han <= 2; break;
                case 100: listed = han >= 1 && han <= 2; break;
                case 110: listed = han >= 1 && han <= 2; break;
                default: listed = false;
            }
            if (!listed) bad();
            base = std::min<i32>((i32)fu << (2 + han), 2000);
        }
    }
    auto pts = [&](i32 mult) { return (base * mult + 99) / 100 * 100; };
    Point p;
    if (is_oya) {
        p.ron = pts(6);
        p.tsumo_ko = pts(2);
        p.tsumo_oya = 0;
    } else {
        p.ron = pts(4);
        p.tsumo_ko = pts(1);
        p.tsumo_oya = pts(2);
    }
    return p;
}

// point.rs:88-103
Point point_yakuman(bool is_oya, i32 count) {
    Point p;
    if (is_oya) {
        p.ron = 48000 * count;
        p.tsumo_ko = 16000 * count;
        p.tsumo_oya = 0;
    } else {
        p.ron = 32000 * count;
        p.tsumo_ko = 8000 * count;
        p.tsumo_oya = 16000 * count;
    }
    return p;
}

// ---------------------------------------------------------------- rankings.rs:8-22
void rankings(const i32* scores, u8* player_by_rank, u8* rank_by_player) {
    u8 pbr[4] = {0, 1, 2, 3};
    std::stable_sort(pbr, pbr + 4, [&](u8 a, u8 b) { return -scores[a] < -scores[b]; });
    for (int r = 0; r < 4; r++) {
        if (player_by_rank) player_by_rank[r] = pbr[r];
        rank_by_player[pbr[r]] = (u8)r;
    }
}

// ---------------------------------------------------------------- agari.rs
// agari.rs:175-190
int agari_cmp(const Agari& l, const Agari& r) {
    if (l.is_yakuman && r.is_yakuman) return (int)l.yakuman - (int)r.yakuman;
    if (l.is_yakuman) return 1;
    if (r.is_yakuman) return -1;
    if (l.han != r.han) return (int)l.han - (int)r.han;
    return (int)l.fu - (int)r.fu;
}

// agari.rs:767-838
u32 get_tile14_and_key(const u8* tiles, u8* tile14) {
    memset(tile14, 0, 14);
    int n14 = 0;
    u32 key = 0;
    int bit_idx = -1;
    bool prev_in_hand = false;
    for (int kind = 0; kind < 3; kind++) {
        for (int num = 0; num < 9; num++) {
            u8 c = tiles[kind * 9 + num];
            if (c > 0) {
                prev_in_hand = true;
                tile14[n14++] = (u8)(kind * 9 + num);
                bit_idx += 1;
                switch (c) {
                    case 2: key |= 0b11u << bit_idx; bit_idx += 2; break;
                    case 3: key |= 0b1111u << bit_idx; bit_idx += 4; break;
                    case 4: key |= 0b111111u << bit_idx; bit_idx += 6; break;
                    default: break;
                }
            } else if (prev_in_hand) {
                prev_in_hand = false;
                key |= 1u << bit_idx;
                bit_idx += 1;
            }
        }
        if (prev_in_hand) {
            prev_in_hand = false;
            key |= 1u << bit_idx;
            bit_idx += 1;
        }
    }
    for (int t = 27; t < 34; t++) {
        u8 c = tiles[t];
        if (c == 0) continue;
        tile14[n14++] = (u8)t;
        bit_idx += 1;
        switch (c) {
            case 2: key |= 0b11u << bit_idx; bit_idx += 2; break;
            case 3: key |= 0b1111u << bit_idx; bit_idx += 4; break;
            case 4: key |= 0b111111u << bit_idx; bit_idx += 6; break;
            default: break;
        }
        key |= 1u << bit_idx;
        bit_idx += 1;
    }
    return key;
}

namespace {

// agari.rs:53-64, 126-157
struct Div {
    u8 pair_idx;
    u8 kotsu_idxs[4]; int n_kotsu;
    u8 shuntsu_idxs[4]; int n_shuntsu;
    bool has_chitoi, has_chuuren, has_ittsuu, has_ryanpeikou, has_ipeikou;
    explicit Div(u32 v) {
        pair_idx = (v >> 6) & 0xF;
        n_kotsu = v & 7;
        for (int i = 0; i < n_kotsu; i++) kotsu_idxs[i] = (v >> (10 + i * 4)) & 0xF;
        n_shuntsu = (v >> 3) & 7;
        for (int i = 0; i < n_shuntsu; i++) shuntsu_idxs[i] = (v >> (10 + (n_kotsu + i) * 4)) & 0xF;
        has_chitoi = (v >> 26) & 1;
        has_chuuren = (v >> 27) & 1;
        has_ittsuu = (v >> 28) & 1;
        has_ryanpeikou = (v >> 29) & 1;
        has_ipeikou = (v >> 30) & 1;
    }
};

inline bool is_sangen(u8 t) { return t == T_P || t == T_F || t == T_C; }
inline bool is_wind(u8 t) { return t >= T_E && t <= T_N; }

// agari.rs:103-124, 287-761
struct DivWorker {
    const AgariCalc& sup;
    const u8* tile14;
    const Div& div;
    u8 pair_tile;
    u8 menzen_kotsu[4]; int n_mk;
    u8 menzen_shuntsu[4]; int n_ms;
    bool winning_tile_makes_minkou;

    // agari.rs:288-312
    DivWorker(const AgariCalc& c, const u8* t14, const Div& d) : sup(c), tile14(t14), div(d) {
        pair_tile = tile14[div.pair_idx];
        n_mk = div.n_kotsu;
        for (int i = 0; i < n_mk; i++) menzen_kotsu[i] = tile14[div.kotsu_idxs[i]];
        n_ms = div.n_shuntsu;
        for (int i = 0; i < n_ms; i++) menzen_shuntsu[i] = tile14[div.shuntsu_idxs[i]];
        winning_tile_makes_minkou = calc_wtmm();
    }

    bool mk_contains(u8 t) const {
        for (int i = 0; i < n_mk; i++) if (menzen_kotsu[i] == t) return true;
        return false;
    }
    bool ms_contains(u8 t) const {
        for (int i = 0; i < n_ms; i++) if (menzen_shuntsu[i] == t) return true;
        return false;
    }

    // agari.rs:315-338
    bool calc_wtmm() const {
        if (!sup.is_ron) return false;
        if (!mk_contains(sup.winning_tile)) return false;
        if (sup.winning_tile >= 27) return true;
        u8 kind = sup.winning_tile / 9, num = sup.winning_tile % 9;
        u8 low = kind * 9 + (num >= 2 ? num - 2 : 0);
        u8 high = kind * 9 + std::min<u8>(num, 6);
        for (u8 t = low; t <= high; t++) if (ms_contains(t)) return false;
        return true;
    }

    // iterators (agari.rs:341-361) materialised into small arrays
    int all_kotsu_and_kantsu(u8* out) const {
        int n = 0;
        for (int i = 0; i < n_mk; i++) out[n++] = menzen_kotsu[i];
        for (int i = 0; i < sup.n_pons; i++) out[n++] = sup.pons[i];
        for (int i = 0; i < sup.n_minkans; i++) out[n++] = sup.minkans[i];
        for (int i = 0; i < sup.n_ankans; i++) out[n++] = sup.ankans[i];
        return n;
    }
    int all_shuntsu(u8* out) const {
        int n = 0;
        for (int i = 0; i < n_ms; i++) out[n++] = menzen_shuntsu[i];
        for (int i = 0; i < sup.n_chis; i++) out[n++] = sup.chis[i];
        return n;
    }

    // agari.rs:362-450
    u8 calc_fu(bool has_pinfu) const {
        if (div.has_chitoi) return 25;
        int fu = 20;
        for (int i = 0; i < n_mk; i++) {
            u8 t = menzen_kotsu[i];
            bool is_minkou = winning_tile_makes_minkou && t == sup.winning_tile;
            bool yao = is_yaokyuu(t);
            if (!is_minkou && yao) fu += 8;
            else if ((!is_minkou && !yao) || (is_minkou && yao)) fu += 4;
            else fu += 2;
        }
        for (int i = 0; i < sup.n_pons; i++) fu += is_yaokyuu(sup.pons[i]) ? 4 : 2;
        for (int i = 0; i < sup.n_ankans; i++) fu += is_yaokyuu(sup.ankans[i]) ? 32 : 16;
        for (int i = 0; i < sup.n_minkans; i++) fu += is_yaokyuu(sup.minkans[i]) ? 16 : 8;

        if (is_sangen(pair_tile)) {
            fu += 2;
        } else {
            if (pair_tile == sup.bakaze) fu += 2;
            if (pair_tile == sup.jikaze) fu += 2;
        }

        if (fu == 20) {
            if (!sup.is_menzen) return 30;
            if (has_pinfu) return sup.is_ron ? 30 : 20;
            return sup.is_ron ? 40 : 30;
        }

        if (!sup.is_ron) fu += 2;
        else if (sup.is_menzen) fu += 10;

        if (!winning_tile_makes_minkou) {
            if (pair_tile == sup.winning_tile) {
                fu += 2;
            } else {
                bool kp = false;
                for (int i = 0; i < n_ms; i++) {
                    u8 s = menzen_shuntsu[i];
                    if (s + 1 == sup.winning_tile || (s % 9 == 0 && s + 2 == sup.winning_tile) ||
                        (s % 9 == 6 && s == sup.winning_tile))
                        kp = true;
                }
                if (kp) fu += 2;
            }
        }
        return (u8)(((fu - 1) / 10 + 1) * 10);
    }

    // agari.rs:452-761
    Agari search_yakus(bool return_if_any) const {
        int han = 0, yakuman = 0;

        bool has_pinfu = n_ms == 4 && !is_sangen(pair_tile) && pair_tile != sup.bakaze && pair_tile != sup.jikaze;
        if (has_pinfu) {
            bool any = false;
            for (int i = 0; i < n_ms; i++) {
                u8 s = menzen_shuntsu[i];
                u8 num = s % 9 + 1;
                if ((num <= 6 && s == sup.winning_tile) || (num >= 2 && s + 2 == sup.winning_tile)) any = true;
            }
            has_pinfu = any;
        }

        auto make_return = [&]() {
            Agari a;
            if (yakuman > 0) {
                a.valid = true; a.is_yakuman = true; a.yakuman = (u8)yakuman;
            } else if (han > 0) {
                a.valid = true;
                a.han = (u8)han;
                a.fu = (return_if_any || han >= 5) ? 0 : calc_fu(has_pinfu);
            }
            return a;
        };
#define CHECK_EARLY_RETURN(stmt) do { stmt; if (return_if_any) return make_return(); } while (0)

        if (has_pinfu) CHECK_EARLY_RETURN(han += 1);
        if (div.has_chitoi) CHECK_EARLY_RETURN(han += 2);
        if (div.has_ryanpeikou) CHECK_EARLY_RETURN(han += 3);
        if (div.has_chuuren) CHECK_EARLY_RETURN(yakuman += 1);

        u8 kk[16]; int n_kk = all_kotsu_and_kantsu(kk);
        u8 ss[8]; int n_ss = all_shuntsu(ss);

        auto is_tanyao_tile = [](u8 k) { u8 kind = k / 9, num = k % 9; return kind < 3 && num > 0 && num < 8; };
        bool has_tanyao;
        if (div.has_chitoi) {
            has_tanyao = true;
            for (int i = 0; i < 7; i++) if (!is_tanyao_tile(tile14[i])) has_tanyao = false;
        } else {
            has_tanyao = true;
            for (int i = 0; i < n_ss; i++) { u8 num = ss[i] % 9; if (!(num > 0 && num < 6)) has_tanyao = false; }
            for (int i = 0; i < n_kk; i++) if (!is_tanyao_tile(kk[i])) has_tanyao = false;
            if (!is_tanyao_tile(pair_tile)) has_tanyao = false;
        }
        if (has_tanyao) CHECK_EARLY_RETURN(han += 1);

        bool has_toitoi = !div.has_chitoi && n_ms == 0 && sup.n_chis == 0;
        if (has_toitoi) CHECK_EARLY_RETURN(han += 2);

        // agari.rs:534-572 — isou scan with take_while semantics
        int isou_kind = -1;
        bool has_jihai = false;
        bool is_chinitsu_or_honitsu = true;
        {
            u8 seq[24]; int n = 0;
            if (div.has_chitoi) {
                for (int i = 0; i < 7; i++) seq[n++] = tile14[i];
            } else {
                for (int i = 0; i < n_kk; i++) seq[n++] = kk[i];
                for (int i = 0; i < n_ss; i++) seq[n++] = ss[i];
                seq[n++] = pair_tile;
            }
            for (int i = 0; i < n; i++) {
                u8 kind = seq[i] / 9;
                if (kind >= 3) { has_jihai = true; continue; }
                if (isou_kind >= 0) {
                    if (isou_kind != kind) { is_chinitsu_or_honitsu = false; break; }
                } else {
                    isou_kind = kind;
                }
            }
        }
        if (isou_kind < 0) {
            CHECK_EARLY_RETURN(yakuman += 1);  // tsuuiisou
        } else if (is_chinitsu_or_honitsu) {
            int n = (has_jihai ? 2 : 5) + (sup.is_menzen ? 1 : 0);
            CHECK_EARLY_RETURN(han += n);
        }

        if (!div.has_chitoi) {
            // ipeikou — agari.rs:574-597
            if (div.has_ipeikou) {
                CHECK_EARLY_RETURN(han += 1);
            } else if (sup.n_ankans > 0 && sup.is_menzen && n_ms >= 2) {
                u8 marks[3] = {0, 0, 0};
                bool ip = false;
                for (int i = 0; i < n_ms; i++) {
                    u8 t = menzen_shuntsu[i];
                    u8 kind = t / 9, num = t % 9;
                    if ((marks[kind] >> num) & 1) { ip = true; break; }
                    marks[kind] |= 1 << num;
                }
                if (ip) CHECK_EARLY_RETURN(han += 1);
            }

            // ittsuu — agari.rs:599-620
            if (sup.is_menzen && div.has_ittsuu) {
                CHECK_EARLY_RETURN(han += 2);
            } else if (sup.n_chis == 0 && div.has_ittsuu) {
                CHECK_EARLY_RETURN(han += 1);
            } else if (n_ms + sup.n_chis >= 3) {
                int kinds[3] = {0, 0, 0};
                for (int i = 0; i < n_ss; i++) {
                    u8 kind = ss[i] / 9, num = ss[i] % 9;
                    if (num == 0) kinds[kind] |= 1;
                    else if (num == 3) kinds[kind] |= 2;
                    else if (num == 6) kinds[kind] |= 4;
                }
                if (kinds[0] == 7 || kinds[1] == 7 || kinds[2] == 7) CHECK_EARLY_RETURN(han += 1);
            }

            // sanshoku — agari.rs:622-647
            int s_counter[9] = {0};
            for (int i = 0; i < n_ss; i++) s_counter[ss[i] % 9] |= 1 << (ss[i] / 9);
            bool doujun = false;
            for (int i = 0; i < 9; i++) if (s_counter[i] == 7) doujun = true;
            if (doujun) {
                int n = sup.is_menzen ? 2 : 1;
                CHECK_EARLY_RETURN(han += n);
            } else {
                int k_counter[9] = {0};
                for (int i = 0; i < n_kk; i++) if (kk[i] / 9 < 3) k_counter[kk[i] % 9] |= 1 << (kk[i] / 9);
                bool doukou = false;
                for (int i = 0; i < 9; i++) if (k_counter[i] == 7) doukou = true;
                if (doukou) CHECK_EARLY_RETURN(han += 2);
            }

            // agari.rs:649-667
            int ankous = sup.n_ankans + n_mk - (winning_tile_makes_minkou ? 1 : 0);
            if (ankous == 4) CHECK_EARLY_RETURN(yakuman += 1);
            else if (ankous == 3) CHECK_EARLY_RETURN(han += 2);

            int kans = sup.n_ankans + sup.n_minkans;
            if (kans == 4) CHECK_EARLY_RETURN(yakuman += 1);
            else if (kans == 3) CHECK_EARLY_RETURN(han += 2);

            // ryuisou — agari.rs:669-677
            auto green = [](u8 k) { return k == T_2S || k == T_3S || k == T_4S || k == T_6S || k == T_8S || k == T_F; };
            bool ryu = green(pair_tile);
            for (int i = 0; i < n_kk; i++) if (!green(kk[i])) ryu = false;
            for (int i = 0; i < n_ss; i++) if (ss[i] != T_2S) ryu = false;
            if (ryu) CHECK_EARLY_RETURN(yakuman += 1);

            // yakuhai etc — agari.rs:679-721
            if (!has_tanyao) {
                bool hj[7] = {false};
                for (int i = 0; i < n_kk; i++) if (kk[i] >= 27) hj[kk[i] - 27] = true;
                if (hj[sup.bakaze - 27]) CHECK_EARLY_RETURN(han += 1);
                if (hj[sup.jikaze - 27]) CHECK_EARLY_RETURN(han += 1);
                int saneins = (int)hj[4] + (int)hj[5] + (int)hj[6];
                if (saneins > 0) {
                    CHECK_EARLY_RETURN(han += saneins);
                    if (saneins == 3) CHECK_EARLY_RETURN(yakuman += 1);
                    else if (saneins == 2 && is_sangen(pair_tile)) CHECK_EARLY_RETURN(han += 2);
                }
                int winds = (int)hj[0] + (int)hj[1] + (int)hj[2] + (int)hj[3];
                if (winds == 4) CHECK_EARLY_RETURN(yakuman += 1);
                else if (winds == 3 && is_wind(pair_tile)) CHECK_EARLY_RETURN(yakuman += 1);
            }
        }

        // agari.rs:724-761
        if (!has_tanyao) {
            bool hj = false;
            // NOTE: `.all()` short-circuits, so has_jihai only reflects elements visited
            // before the first non-yaokyuu element; when the result is true all were visited.
            auto is_yk = [&](u8 k) {
                u8 kind = k / 9;
                if (kind >= 3) { hj = true; return true; }
                u8 num = k % 9;
                return num == 0 || num == 8;
            };
            bool all_yk = true;
            if (div.has_chitoi) {
                for (int i = 0; i < 7 && all_yk; i++) if (!is_yk(tile14[i])) all_yk = false;
            } else {
                for (int i = 0; i < n_kk && all_yk; i++) if (!is_yk(kk[i])) all_yk = false;
                if (all_yk && !is_yk(pair_tile)) all_yk = false;
            }
            if (all_yk) {
                if (div.has_chitoi || has_toitoi) {
                    if (hj) CHECK_EARLY_RETURN(han += 2);  // honroutou
                    else CHECK_EARLY_RETURN(yakuman += 1);  // chinroutou
                } else {
                    bool jc = true;
                    for (int i = 0; i < n_ss; i++) { u8 num = ss[i] % 9; if (!(num == 0 || num == 6)) jc = false; }
                    if (jc) {
                        int n = (hj ? 1 : 2) + (sup.is_menzen ? 1 : 0);
                        CHECK_EARLY_RETURN(han += n);
                    }
                }
            }
        }
#undef CHECK_EARLY_RETURN
        return make_return();
    }
};

}  // namespace

// agari.rs:257-285
static Agari search_yakus_impl(const AgariCalc& c, bool return_if_any) {
    bool menzen_expected = c.n_chis == 0 && c.n_pons == 0 && c.n_minkans == 0;
    if (c.is_menzen != menzen_expected) throw OrcError("agari: is_menzen inconsistent with melds");

    if (c.is_menzen && shanten_kokushi(c.tehai) == -1) {
        Agari a; a.valid = true; a.is_yakuman = true; a.yakuman = 1;
        return a;
    }
    u8 tile14[14];
    u32 key = get_tile14_and_key(c.tehai, tile14);
    u32 divs[4];
    int n = agari_table_lookup(key, divs);
    if (n < 0) return Agari{};

    Agari best;
    for (int i = 0; i < n; i++) {
        Div d(divs[i]);
        DivWorker w(c, tile14, d);
        Agari a = w.search_yakus(return_if_any);
        if (!a.valid) continue;
        if (return_if_any) return a;
        if (!best.valid || agari_cmp(a, best) >= 0) best = a;
    }
    return best;
}

bool AgariCalc::has_yaku() const { return search_yakus_impl(*this, true).valid; }
Agari AgariCalc::search_yakus() const { return search_yakus_impl(*this, false); }

// agari.rs:225-255
Agari AgariCalc::agari(u8 additional_hans, u8 doras) const {
    Agari a = search_yakus();
    if (a.valid) {
        if (!a.is_yakuman) a.han = (u8)(a.han + additional_hans + doras);
        return a;
    }
    if (additional_hans == 0) return Agari{};
    if (additional_hans + doras >= 5) {
        Agari r; r.valid = true; r.fu = 0; r.han = (u8)(additional_hans + doras);
        return r;
    }
    u8 tile14[14];
    u32 key = get_tile14_and_key(tehai, tile14);
    u32 divs[4];
    int n = agari_table_lookup(key, divs);
    if (n <= 0) return Agari{};
    u8 fu = 0;
    for (int i = 0; i < n; i++) {
        Div d(divs[i]);
        DivWorker w(*this, tile14, d);
        fu = std::max(fu, w.calc_fu(false));
    }
    Agari r; r.valid = true; r.fu = fu; r.han = (u8)(additional_hans + doras);
    return r;
}

// agari.rs:854-912
bool check_ankan_after_riichi(const u8* tehai, u8 len_div3, u8 tile, bool strict) {
    int tile_id = deaka(tile);
    if (tehai[tile_id] != 4) return false;
    if (tile_id >= 27) return true;

    u8 before_tsumo[34];
    memcpy(before_tsumo, tehai, 34);
    before_tsumo[tile_id] -= 1;

    for (int t = 0; t < 34; t++) {
        if (before_tsumo[t] == 4) continue;
        u8 tmp[34];
        memcpy(tmp, before_tsumo, 34);
        tmp[t] += 1;
        if (shanten_all(tmp, len_div3) != -1) continue;
        int wait = t;
        if (wait == tile_id) return false;
        u8 after[34];
        memcpy(after, tehai, 34);
        after[tile_id] = 0;
        after[wait] += 1;
        u8 t14[14];
        u32 divs_after[4];
        int n_after = agari_table_lookup(get_tile14_and_key(after, t14), divs_after);
        if (n_after < 0) return false;
        if (strict) {
            u8 before[34];
            memcpy(before, before_tsumo, 34);
            before[wait] += 1;
            u32 divs_before[4];
            int n_before = agari_table_lookup(get_tile14_and_key(before, t14), divs_before);
            if (n_before < 0) throw OrcError("invalid riichi detected when testing ankan after riichi");
            if (n_after != n_before) return false;
        }
    }
    return true;
}

}  // namespace orc
