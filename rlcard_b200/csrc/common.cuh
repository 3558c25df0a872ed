// common.cuh -- shared device utilities: Philox/tape/MT19937 chance sources, Fisher-Yates,
// packed-state helpers and the warp-cooperative coalesced row writer.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rlc {

constexpr int kWarp = 32;

// ------------------------------------------------------------------------------------------
// Philox4x32-10, throughput-mode draw layout (DESIGN.md "Philox streams").
// One block per env-step, computed unconditionally (no divergence between lanes whose episodes end
// at different times):  block(b) of step (episode e, step-in-episode t) of env i has
//     counter = (t, e, global env id, b),  key = 64-bit seed.
//   block 0 = [ policy word | chance draw 0 | chance draw 1 | chance draw 2 ]
//   block b>=1 = chance draws 3+4(b-1) .. 6+4(b-1)
// "chance draws of a step" are, in order, every bounded draw the engine makes while applying the
// action of that step AND while dealing the next episode if the step ended one (auto reset).
// A reset that is not part of a step (rlc_reset) uses t = 0xffffffff of the episode it starts.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t &o0, uint32_t &o1,
                                              uint32_t &o2, uint32_t &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
constexpr uint32_t kResetStep = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// Chance sources.  All engines draw through below(n) (uniform in [0,n)); skip_fy(i_hi, i_lo)
// stands for the Fisher-Yates steps i = i_hi .. i_lo whose outcome no kernel ever observes:
// Philox mode does not make them, tape/MT modes must consume them to stay aligned with the
// reference's np.random stream.
// ------------------------------------------------------------------------------------------
struct ChancePhilox {
    static constexpr int kKind = 0;
    uint32_t k0, k1, env, episode, t;
    uint32_t w0, w1, w2, w3;           // block 0 of the current step
    uint32_t x0, x1, x2, x3, xblk;     // one cached extra block
    uint32_t d; int err;
    __device__ __forceinline__ void init(uint64_t seed, uint32_t env_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); env = env_; err = 0; d = 0; xblk = 0;
        w0 = w1 = w2 = w3 = x0 = x1 = x2 = x3 = 0; episode = t = 0;
    }
    __device__ __forceinline__ void begin(uint32_t episode_, uint32_t t_) {
        episode = episode_; t = t_; d = 0; xblk = 0;
        philox4x32_10(t, episode, env, 0u, k0, k1, w0, w1, w2, w3);
    }
    __device__ __forceinline__ uint32_t policy_word() const { return w0; }
    __device__ __forceinline__ uint32_t raw() {
        uint32_t r;
        if (d < 3u) r = d == 0 ? w1 : (d == 1 ? w2 : w3);
        else {
            const uint32_t q = 1u + ((d - 3u) >> 2), l = (d - 3u) & 3u;
            if (q != xblk) { philox4x32_10(t, episode, env, q, k0, k1, x0, x1, x2, x3); xblk = q; }
            const uint32_t lo = (l & 1u) ? x1 : x0, hi = (l & 1u) ? x3 : x2;
            r = (l & 2u) ? hi : lo;
        }
        d++;
        return r;
    }
    __device__ __forceinline__ uint32_t below(uint32_t n) { return __umulhi(raw(), n); }
    __device__ __forceinline__ void skip_fy(int, int) {}
};

struct ChanceTape {
    static constexpr int kKind = 1;
    const uint8_t *tape; int len, pos, err;
    __device__ __forceinline__ uint32_t below(uint32_t n) {
        uint32_t v = 0;
        if (pos >= len) err |= 1; else v = tape[pos];
        pos++;
        if (v >= n) { err |= 2; v = n - 1; }
        return v;
    }
    __device__ __forceinline__ void skip_fy(int i_hi, int i_lo) { if (i_hi >= i_lo) pos += i_hi - i_lo + 1; }
    __device__ __forceinline__ void begin(uint32_t, uint32_t) {}
};

// np.random.RandomState (MT19937) with numpy's legacy bounded draws: mask-and-reject on 32-bit
// words, no word consumed when the range is a single value.  State: 624 words + index, one
// column per env ([625][n]).
struct ChanceMt {
    static constexpr int kKind = 2;
    uint32_t *mt; size_t n; int mti, err;
    __device__ __forceinline__ uint32_t &w(int j) { return mt[(size_t)j * n]; }
    __device__ void twist() {
        for (int kk = 0; kk < 624; kk++) {
            const uint32_t y = (w(kk) & 0x80000000u) | (w(kk == 623 ? 0 : kk + 1) & 0x7fffffffu);
            const int m = kk < 227 ? kk + 397 : kk - 227;
            w(kk) = w(m) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        mti = 0;
    }
    __device__ uint32_t next() {
        if (mti >= 624) twist();
        uint32_t y = w(mti++);
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    __device__ uint32_t below(uint32_t nn) {
        const uint32_t mx = nn - 1;
        if (mx == 0) return 0;
        uint32_t mask = mx;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next() & mask; } while (v > mx);
        return v;
    }
    __device__ void skip_fy(int i_hi, int i_lo) { for (int i = i_hi; i >= i_lo; i--) (void)below((uint32_t)i + 1u); }
    __device__ __forceinline__ void begin(uint32_t, uint32_t) {}
};

// RandomState.shuffle on a list: for i = n-1 .. 1: j = below(i+1); swap(x[i], x[j]).
// Only the last `tail` positions are observed by the caller (deck.pop() deals from the end).
template <class Ch>
__device__ __forceinline__ void shuffle_tail_u8(Ch &ch, uint8_t *x, int n, int tail) {
    const int stop = (tail >= n - 1) ? 1 : n - tail;
    for (int i = n - 1; i >= stop; i--) {
        const uint32_t j = ch.below((uint32_t)i + 1u);
        const uint8_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
    ch.skip_fy(stop - 1, 1);
}

// ------------------------------------------------------------------------------------------
// bit packing helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf_get(uint32_t w, int pos, int len) { return (w >> pos) & ((1u << len) - 1u); }
__device__ __forceinline__ uint32_t bf_put(uint32_t v, int pos) { return v << pos; }

// k-th (0-based) set bit of a multi-word mask
template <int WORDS>
__device__ __forceinline__ int kth_set_bit(const uint32_t (&m)[WORDS], int k) {
#pragma unroll
    for (int w = 0; w < WORDS; w++) {
        const int c = __popc(m[w]);
        if (k < c) return 32 * w + (int)__fns(m[w], 0, k + 1);
        k -= c;
    }
    return -1;
}
template <int WORDS>
__device__ __forceinline__ int popc_words(const uint32_t (&m)[WORDS]) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < WORDS; w++) c += __popc(m[w]);
    return c;
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative row writer.  Each lane fills its own row of a per-warp shared-memory tile
// (rows of ROW_BYTES, one env per lane); the 32 rows are contiguous in global memory, so the
// warp then streams the tile out with 128-bit stores (st.global.cs: written once, never re-read
// by this kernel) and re-zeroes the tile on the way.  Falls back to byte stores when the
// destination or the length is not 16-byte aligned (ragged last warp).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_zero(uint8_t *tile, int nbytes, int lane) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = lane; c < (nbytes >> 4); c += kWarp) reinterpret_cast<uint4 *>(tile)[c] = z;
    for (int c = (nbytes & ~15) + lane; c < nbytes; c += kWarp) tile[c] = 0;
}
__device__ __forceinline__ void warp_tile_flush(uint8_t *gdst, uint8_t *tile, int nbytes, int lane) {
    if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)nbytes) & 15u) == 0) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int c = lane; c < (nbytes >> 4); c += kWarp) {
            const uint4 v = reinterpret_cast<uint4 *>(tile)[c];
            __stcs(reinterpret_cast<uint4 *>(gdst) + c, v);
            reinterpret_cast<uint4 *>(tile)[c] = z;
        }
    } else {
        for (int c = lane; c < nbytes; c += kWarp) { gdst[c] = tile[c]; tile[c] = 0; }
    }
}

// per-env header words stored in front of the game words of the state
struct EnvHeader {
    uint32_t episode;   // episodes started (0 = never reset); the running episode has index episode-1
    uint32_t t;         // env-steps taken in the running episode
    __device__ __forceinline__ void load(const uint32_t *st, size_t n, size_t i) { episode = st[i]; t = st[n + i]; }
    __device__ __forceinline__ void store(uint32_t *st, size_t n, size_t i) const { st[i] = episode; st[n + i] = t; }
};
constexpr int kHeaderWords = 2;

}  // namespace rlc
