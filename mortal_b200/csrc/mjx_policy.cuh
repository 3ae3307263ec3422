// mortal_b200 — counter-based TEST policies (not part of the reference): used by the parity tests and
// the env-only benchmark so that the CUDA environment and the CPU oracle can be driven by the very
// same decisions without exchanging data. Definition shared with oracle/board.cc::test_policy.
#pragma once
#include "mjx_types.cuh"

namespace mjx {

MJX_HD u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

MJX_HD u64 policy_hash(u64 nonce, u64 key, u64 table, u64 step_idx, u32 seat, u32 kan) {
    u64 h = splitmix64(nonce);
    h = splitmix64(h ^ key);
    h = splitmix64(h ^ table);
    h = splitmix64(h ^ step_idx);
    h = splitmix64(h ^ (u64)(seat * 2 + kan));
    return h;
}

// index of the k-th set bit of m (k < popcount(m))
MJX_D int kth_set_bit(u64 m, int k) {
    for (int i = 0; i < k; i++) m &= m - 1;
    return mjx_ffsll(m) - 1;
}

// kind 0: uniform over the legal mask. kind 1: agari first, riichi with p = 3/4, calls vs discards
// by coin flip, discards prefer next-shanten then keep-shanten tiles (see oracle/board.cc).
// kind 2: kind 1 with the hash taken from the legal mask alone (a function of what an engine sees; bench.py's e2e engine).
MJX_D int test_policy(int kind, u64 h, bool kan_select, u64 mask, u64 keep34, u64 next34) {
    if (kind == 2) { h = splitmix64(mask); kind = 1; }
    if (kind == 0 || kan_select) return kth_set_bit(mask, (int)(h % (u64)mjx_popcll(mask)));
    if ((mask >> 43) & 1) return 43;
    u64 h2 = splitmix64(h);
    if (((mask >> 37) & 1) && (h2 & 3) != 0) return 37;
    const u64 DISC = (1ull << 37) - 1;
    u64 disc = mask & DISC;
    u64 other = mask & ~DISC & ~(1ull << 37);
    int n_disc = mjx_popcll(disc), n_other = mjx_popcll(other);
    u64 h3 = splitmix64(h2);
    if (n_other > 0 && (n_disc == 0 || (h3 & 1))) return kth_set_bit(other, (int)((h3 >> 1) % (u64)n_other));
    // planes are indexed by the de-aka'd tile: spread 5m/5p/5s bits onto the aka ids 34..36
    u64 ak = ((next34 >> 4) & 1) << 34 | ((next34 >> 13) & 1) << 35 | ((next34 >> 22) & 1) << 36;
    u64 pref = disc & (next34 | ak);
    if (!pref) {
        ak = ((keep34 >> 4) & 1) << 34 | ((keep34 >> 13) & 1) << 35 | ((keep34 >> 22) & 1) << 36;
        pref = disc & (keep34 | ak);
    }
    if (!pref) pref = disc;
    return kth_set_bit(pref, (int)((h3 >> 1) % (u64)mjx_popcll(pref)));
}

}  // namespace mjx
