// mortal_b200 — per-kyoku wall generation on device.
// Contract: arena/board.rs:99-123 + 786-824: seed32 = SHA3-256(nonce_le8 || key_le8 || [kyoku, honba]),
// ChaCha12 keystream (64-bit counter from 0), shuffle of the 136-tile UNSHUFFLED sequence.
// shuffle_kind 1 = rand 0.8 (verified against the reference's seeded log), 0 = rand 0.9.1 restatement
// (nominal for libriichi@d5e80bf, unpinned by any reference fixture — see DESIGN.md).
#pragma once
#include "mjx_types.cuh"

namespace mjx {

MJX_CONST u64 c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

MJX_D u64 rotl64(u64 x, int n) { return (x << n) | (x >> ((64 - n) & 63)); }
MJX_D u32 rotl32(u32 x, int n) { return mjx_rotl32(x, n); }

// Single-block SHA3-256 of an 18-byte message; out = 8 little-endian u32 words of the digest.
MJX_DN void sha3_256_seed(u64 nonce, u64 key, u32 kyoku, u32 honba, u32* out8) {
    u64 s[25];
#pragma unroll
    for (int i = 0; i < 25; i++) s[i] = 0;
    s[0] = nonce;
    s[1] = key;
    // bytes 16,17 = kyoku, honba; byte 18 = 0x06 domain pad; byte 135 (lane 16, top byte) = 0x80
    s[2] = (u64)(kyoku & 0xFF) | ((u64)(honba & 0xFF) << 8) | (0x06ull << 16);
    s[16] = 0x8000000000000000ULL;
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        u64 C[5], B[25];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            u64 D = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
#pragma unroll
            for (int y = 0; y < 5; y++) s[x + 5 * y] ^= D;
        }
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(s[x + 5 * y], ROT[x + 5 * y]);
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= c_keccak_rc[round];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out8[2 * i] = (u32)s[i];
        out8[2 * i + 1] = (u32)(s[i] >> 32);
    }
}

struct ChaCha12 {
    u32 key[8];
    u32 counter;
    u32 buf[16];
    int pos;
};

MJX_DN void chacha12_block(ChaCha12& r) {
    u32 x[16];
    x[0] = 0x61707865; x[1] = 0x3320646e; x[2] = 0x79622d32; x[3] = 0x6b206574;
#pragma unroll
    for (int i = 0; i < 8; i++) x[4 + i] = r.key[i];
    x[12] = r.counter; x[13] = 0; x[14] = 0; x[15] = 0;
    u32 st[16];
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = x[i];
#define MJX_QR(a, b, c, d)                               \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);        \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);        \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);         \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        MJX_QR(0, 4, 8, 12) MJX_QR(1, 5, 9, 13) MJX_QR(2, 6, 10, 14) MJX_QR(3, 7, 11, 15)
        MJX_QR(0, 5, 10, 15) MJX_QR(1, 6, 11, 12) MJX_QR(2, 7, 8, 13) MJX_QR(3, 4, 9, 14)
    }
#undef MJX_QR
#pragma unroll
    for (int i = 0; i < 16; i++) r.buf[i] = x[i] + st[i];
    r.counter++;
    r.pos = 0;
}

MJX_D u32 chacha_next(ChaCha12& r) {
    if (r.pos >= 16) chacha12_block(r);
    // dynamic index into a small local array: the shuffle is a once-per-kyoku scalar tail
    return r.buf[r.pos++];
}

// Writes the shuffled 136-tile wall into `wall` (shared memory). Executed by ONE lane.
MJX_DN void make_wall(u64 nonce, u64 key, int kyoku, int honba, int shuffle_kind, u8* wall) {
    ChaCha12 rng;
    sha3_256_seed(nonce, key, (u32)kyoku, (u32)honba, rng.key);
    rng.counter = 0;
    rng.pos = 16;
    for (int i = 0; i < 136; i++) wall[i] = (u8)(i >> 2);
    wall[T_5M * 4] = T_5MR;
    wall[T_5P * 4] = T_5PR;
    wall[T_5S * 4] = T_5SR;
    if (shuffle_kind == 1) {
        // rand 0.8: for i in (1..n).rev(): swap(i, below(i+1)); widening multiply + zone rejection
        for (int i = 135; i >= 1; i--) {
            u32 range = (u32)i + 1;
            u32 zone = (range << mjx_clz(range)) - 1;
            u32 j;
            for (;;) {
                u32 v = chacha_next(rng);
                u64 m = (u64)v * range;
                if ((u32)m <= zone) { j = (u32)(m >> 32); break; }
            }
            u8 a = wall[i]; wall[i] = wall[j]; wall[j] = a;
        }
    } else {
        // rand 0.9.1: forward Fisher-Yates driven by IncreasingUniform (chunked Canon sampling)
        u32 n = 0, chunk = 0;
        int chunk_remaining = 1;
        for (int i = 0; i < 136; i++) {
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
                u64 m = (u64)chacha_next(rng) * product;
                u32 hi = (u32)(m >> 32), lo = (u32)m;
                if (lo > (0u - product)) {
                    u32 new_hi = (u32)(((u64)chacha_next(rng) * product) >> 32);
                    u32 sum = lo + new_hi;
                    if (sum < lo) hi += 1;
                }
                chunk = hi;
                next_cr = remaining - 1;
            } else {
                next_cr = chunk_remaining - 1;
            }
            u32 result;
            if (next_cr == 0) result = chunk;
            else { result = chunk % next_n; chunk /= next_n; }
            chunk_remaining = next_cr;
            n = next_n;
            u8 a = wall[i]; wall[i] = wall[result]; wall[result] = a;
        }
    }
}

}  // namespace mjx
