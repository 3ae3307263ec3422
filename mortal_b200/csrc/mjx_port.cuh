// mortal_b200 — warp-programming vocabulary used by the step / encode code.
//
// On the device (the product) these expand to lane guards, __syncwarp() and warp ballots.
// With -DMJX_HOST_EMUL (used ONLY by tests/host_emul: a single-lane g++ build of the very same
// sources, so rule logic can be diffed against the oracle in the GPU-less dev container) they
// expand to plain loops. The emulation build is test infrastructure; nothing in mortal_b200/
// loads it and there is no CPU fallback in the product path.
#pragma once
#include <cstdint>

#ifdef MJX_HOST_EMUL
#include <algorithm>
#include <cmath>
#include <cstring>
#define MJX_HD static inline
#define MJX_D static inline
#define MJX_DN static inline
#define MJX_DM inline
#define MJX_CONST static const
#define MJX_LDG(p) (*(p))
#define MJX_SYNCWARP() ((void)0)
namespace mjx {
struct U4 { uint32_t x, y, z, w; };
static inline int mjx_popc(uint32_t v) { return __builtin_popcount(v); }
static inline int mjx_popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int mjx_ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
static inline int mjx_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t mjx_rotl32(uint32_t x, int n) { return (x << n) | (x >> ((32 - n) & 31)); }
using std::max;
using std::min;
}  // namespace mjx
#else
#include <cuda_runtime.h>
#define MJX_HD __host__ __device__ __forceinline__
#define MJX_D __device__ __forceinline__
#define MJX_DN __device__ inline
#define MJX_DM __device__ __forceinline__
#define MJX_CONST __constant__
#define MJX_LDG(p) __ldg(p)
#define MJX_SYNCWARP() __syncwarp()
namespace mjx {
typedef uint4 U4;
__device__ __forceinline__ int mjx_popc(uint32_t v) { return __popc(v); }
__device__ __forceinline__ int mjx_popcll(uint64_t v) { return __popcll(v); }
__device__ __forceinline__ int mjx_ffsll(uint64_t v) { return __ffsll((long long)v); }
__device__ __forceinline__ int mjx_clz(uint32_t v) { return __clz(v); }
__device__ __forceinline__ uint32_t mjx_rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
}  // namespace mjx
#endif

// ---- lane vocabulary (c is a Ctx with .lane) ----
#ifdef MJX_HOST_EMUL
// one lane plays all roles
#define MJX_L0(...) do { __VA_ARGS__; } while (0)
#define MJX_IS_L0(c) (true)
#define MJX_FOR_SEATS(c, s) for (int s = 0; s < 4; s++)
#define MJX_END_SEATS(c) ((void)0)
#define MJX_FOR_TILES(c, t) for (int t = 0; t < 34; t++)
#define MJX_END_TILES(c) ((void)0)
#else
#define MJX_L0(...) do { if (c.lane == 0) { __VA_ARGS__; } __syncwarp(); } while (0)
#define MJX_IS_L0(c) ((c).lane == 0)
// lanes 0..3 <-> seats 0..3, all four seats processed concurrently
#define MJX_FOR_SEATS(c, s) if ((c).lane < 4) for (int s = (c).lane, mjx_once_ = 1; mjx_once_; mjx_once_ = 0)
#define MJX_END_SEATS(c) __syncwarp()
#define MJX_FOR_TILES(c, t) for (int t = (c).lane; t < 34; t += 32)
#define MJX_END_TILES(c) __syncwarp()
#endif
