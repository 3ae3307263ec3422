// mortal_b200 — host-side construction of the device lookup tables from the raw table files
// (formats: libriichi algo/shanten.rs:27-44, algo/agari.rs:24-37).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "mjx_types.cuh"
#include "mjx_algo.cuh"

namespace mjx {

struct HostTables {
    std::vector<u64> suhai, jihai;
    std::vector<u32> agari_keys;
    std::vector<U4> agari_divs;
    std::vector<u8> agari_ndivs;
    std::string error;
};

inline bool read_all(const std::string& path, std::vector<u8>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    u8 buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
    fclose(f);
    return true;
}

// 5 bytes per row, low nibble first -> one u64 per row with nibble i at bits [4i, 4i+4)
inline bool pack_rows(const std::vector<u8>& raw, size_t n_rows, size_t padded, std::vector<u64>& out) {
    if (raw.size() != n_rows * 5) return false;
    out.assign(padded, 0);
    for (size_t r = 0; r < n_rows; r++) {
        u64 v = 0;
        for (int b = 0; b < 5; b++) v |= (u64)raw[r * 5 + b] << (8 * b);
        out[r] = v;
    }
    return true;
}

inline bool load_host_tables(const std::string& dir, HostTables& H) {
    std::vector<u8> raw;
    if (!read_all(dir + "/shanten_suhai.bin", raw) || !pack_rows(raw, 1940777, SUHAI_ROWS, H.suhai)) {
        H.error = "cannot load " + dir + "/shanten_suhai.bin (run tools/build_tables.py)";
        return false;
    }
    raw.clear();
    if (!read_all(dir + "/shanten_jihai.bin", raw) || !pack_rows(raw, 78032, JIHAI_ROWS, H.jihai)) {
        H.error = "cannot load " + dir + "/shanten_jihai.bin";
        return false;
    }
    raw.clear();
    if (!read_all(dir + "/agari.bin", raw)) {
        H.error = "cannot load " + dir + "/agari.bin";
        return false;
    }
    H.agari_keys.assign(AGARI_SLOTS, 0xFFFFFFFFu);
    U4 zero; zero.x = zero.y = zero.z = zero.w = 0;
    H.agari_divs.assign(AGARI_SLOTS, zero);
    H.agari_ndivs.assign(AGARI_SLOTS, 0);
    size_t p = 0;
    auto rd32 = [&]() {
        u32 v = raw[p] | (raw[p + 1] << 8) | (raw[p + 2] << 16) | ((u32)raw[p + 3] << 24);
        p += 4;
        return v;
    };
    for (int i = 0; i < 9362; i++) {
        if (p + 5 > raw.size()) { H.error = "agari.bin truncated"; return false; }
        u32 key = rd32();
        int n = raw[p++];
        if (n > 4 || key == 0xFFFFFFFFu || p + 4 * (size_t)n > raw.size()) { H.error = "agari.bin malformed"; return false; }
        u32 d[4] = {0, 0, 0, 0};
        for (int j = 0; j < n; j++) d[j] = rd32();
        u32 slot = agari_hash(key) & (AGARI_SLOTS - 1);
        while (H.agari_keys[slot] != 0xFFFFFFFFu) {
            if (H.agari_keys[slot] == key) { H.error = "agari.bin duplicate key"; return false; }
            slot = (slot + 1) & (AGARI_SLOTS - 1);
        }
        H.agari_keys[slot] = key;
        U4 v; v.x = d[0]; v.y = d[1]; v.z = d[2]; v.w = d[3];
        H.agari_divs[slot] = v;
        H.agari_ndivs[slot] = (u8)n;
    }
    if (p != raw.size()) { H.error = "agari.bin has trailing bytes"; return false; }
    return true;
}

}  // namespace mjx
