// common.cuh -- shared device utilities: Philox/tape/MT19937 chance sources, Fisher-Yates,
// packed-state helpers and the warp-cooperative coalesced row writer.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rlc {

constexpr int kWarp = 32;

// ------------------------------------------------------------------------------------------
// Philox4x32-10, throughput-mode draw layout (DESIGN.md "Philox streams"; oracle/orc_common.c is
// the CPU twin).  key = 64-bit seed.  Every env counts the env-steps it has taken since creation
// (k, header word 2); all randomness of step k is a pure function of (seed, global env id, k):
//   base word   W_k = Philox(ctr = (k >> 2, 0, env, 0))[k & 3]     one block serves 4 consecutive steps
//   policy      index of the action among the cnt legal ones = mulhi(W_k, cnt); chain register R = lo32(W_k * cnt)
//   chain(n)    v = mulhi(R, n), R = lo32(R * n)                    small draws peeled off the same word
//   below(n)    v = mulhi(F_j, n), F_j = Philox(ctr = (k, j >> 2, env, 1))[j & 3], j = 0, 1, ... per step
// "draws of a step" are, in order, every bounded draw the engine makes while applying the action of
// that step AND while dealing the next episode if the step ended one (auto reset).  A reset that is
// not part of a step (rlc_reset, first deal of a rollout) uses domain 2: F_j = Philox(ctr = (k, j >> 2,
// env, 2))[j & 3]; its chain register starts as R = F_0 and its fresh draws at j = 1.
// chain() is only used where the product of the ranges stays tiny (Leduc: cnt * 120 * 2 <= 960), so the
// non-uniformity of the last value is below 2^-22.
//   deal word   D_j = Philox(ctr = (E, j >> 2, env, 3))[j & 3], E = 1, 2, ... the ordinal of the episode being dealt.
// Games with kEpisodeDeal (Limit Hold'em) draw their whole deal from D_0..D_3: the deal is then a pure function of
// (seed, global env id, episode ordinal) -- not of the step at which the previous episode happened to end -- so the
// fused rollout can have a dealer warp prepare the deals of future episodes while the env warp is still playing.
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
__device__ __forceinline__ uint32_t sel4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t i) {
    const uint32_t lo = (i & 1u) ? b : a, hi = (i & 1u) ? d : c;
    return (i & 2u) ? hi : lo;
}
enum { kDomBase = 0, kDomStep = 1, kDomReset = 2, kDomDeal = 3 };

// ------------------------------------------------------------------------------------------
// Chance sources.  All engines draw through below(n) (uniform in [0,n)); skip_fy(i_hi, i_lo)
// stands for the Fisher-Yates steps i = i_hi .. i_lo whose outcome no kernel ever observes:
// Philox mode does not make them, tape/MT modes must consume them to stay aligned with the
// reference's np.random stream.
// ------------------------------------------------------------------------------------------
struct ChancePhilox {
    static constexpr int kKind = 0;
    uint32_t k0, k1, env, k, dom;
    uint32_t b0, b1, b2, b3, btag;     // cached base block (index k >> 2)
    uint32_t x0, x1, x2, x3, xtag;     // cached fresh block of the current (k, dom)
    uint32_t R, d; int err;
    uint32_t ep;                       // ordinal of the episode being dealt (deal words)
    __device__ __forceinline__ void begin_episode(uint32_t e) { ep = e; }
    __device__ __forceinline__ void deal_words(uint32_t &d0, uint32_t &d1, uint32_t &d2, uint32_t &d3) const {
        philox4x32_10(ep, 0u, env, (uint32_t)kDomDeal, k0, k1, d0, d1, d2, d3);
    }
    __device__ __forceinline__ void init(uint64_t seed, uint32_t env_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); env = env_; err = 0; d = 0; k = 0; dom = kDomStep; R = 0; ep = 0;
        b0 = b1 = b2 = b3 = x0 = x1 = x2 = x3 = 0; btag = xtag = 0xffffffffu;
    }
    // step k begins: returns the base word W_k (policy word)
    __device__ __forceinline__ uint32_t begin_step(uint32_t k_) {
        k = k_; dom = kDomStep; d = 0; xtag = 0xffffffffu;
        const uint32_t q = k_ >> 2;
        if (q != btag) { philox4x32_10(q, 0u, env, (uint32_t)kDomBase, k0, k1, b0, b1, b2, b3); btag = q; }
        return sel4(b0, b1, b2, b3, k_ & 3u);
    }
    // the policy consumed mulhi(W, cnt); what is left of the word feeds chain()
    __device__ __forceinline__ void seed_chain(uint32_t w, uint32_t cnt) { R = w * cnt; }
    // a reset outside a step, at lifetime step count k
    __device__ __forceinline__ void begin_reset(uint32_t k_) {
        k = k_; dom = kDomReset;
        philox4x32_10(k_, 0u, env, (uint32_t)kDomReset, k0, k1, x0, x1, x2, x3); xtag = 0;
        R = x0; d = 1;
    }
    __device__ __forceinline__ uint32_t raw() {
        const uint32_t q = d >> 2, l = d & 3u;
        if (q != xtag) { philox4x32_10(k, q, env, dom, k0, k1, x0, x1, x2, x3); xtag = q; }
        d++;
        return sel4(x0, x1, x2, x3, l);
    }
    __device__ __forceinline__ uint32_t below(uint32_t n) { return __umulhi(raw(), n); }
    __device__ __forceinline__ uint32_t chain(uint32_t n) { const uint32_t v = __umulhi(R, n); R *= n; return v; }
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
    __device__ __forceinline__ uint32_t chain(uint32_t n) { return below(n); }
    __device__ __forceinline__ uint32_t begin_step(uint32_t) { return 0u; }
    __device__ __forceinline__ void seed_chain(uint32_t, uint32_t) {}
    __device__ __forceinline__ void begin_reset(uint32_t) {}
    __device__ __forceinline__ void begin_episode(uint32_t) {}
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
    __device__ __forceinline__ uint32_t chain(uint32_t n) { return below(n); }
    __device__ __forceinline__ uint32_t begin_step(uint32_t) { return 0u; }
    __device__ __forceinline__ void seed_chain(uint32_t, uint32_t) {}
    __device__ __forceinline__ void begin_reset(uint32_t) {}
    __device__ __forceinline__ void begin_episode(uint32_t) {}
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

// position of the n-th (0-based) set bit of m, n < popc(m): five popc halvings, branch free
// (the __fns intrinsic expands to ~50 instructions; this is ~30 and has no sign / offset handling)
__device__ __forceinline__ int nth_set_bit32(uint32_t m, int n) {
    int pos = 0, c;
    c = __popc(m & 0xffffu);          if (n >= c) { n -= c; pos += 16; }
    c = __popc((m >> pos) & 0xffu);   if (n >= c) { n -= c; pos += 8; }
    c = __popc((m >> pos) & 0xfu);    if (n >= c) { n -= c; pos += 4; }
    c = __popc((m >> pos) & 0x3u);    if (n >= c) { n -= c; pos += 2; }
    c = (int)((m >> pos) & 1u);       if (n >= c) { pos += 1; }
    return pos;
}

// k-th (0-based) set bit of a multi-word mask
template <int WORDS>
__device__ __forceinline__ int kth_set_bit(const uint32_t (&m)[WORDS], int k) {
#pragma unroll
    for (int w = 0; w < WORDS; w++) {
        const int c = __popc(m[w]);
        if (k < c) return 32 * w + nth_set_bit32(m[w], k);
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

// Trajectory / obs / mask rows are written once and never re-read by the kernels: streaming stores.  RLC_STORE_POLICY
// selects the cache operator for tuning runs (0 = st.global.cs, the default; 1 = plain; 2 = .wt; 3 = .cg).
#ifndef RLC_STORE_POLICY
#define RLC_STORE_POLICY 0
#endif
template <class T>
__device__ __forceinline__ void st_stream(T *p, T v) {
#if RLC_STORE_POLICY == 0
    __stcs(p, v);
#elif RLC_STORE_POLICY == 1
    *p = v;
#elif RLC_STORE_POLICY == 2
    __stwt(p, v);
#else
    __stcg(p, v);
#endif
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative row writer.  Each lane fills its own row of a per-warp shared-memory tile
// (rows of ROW_BYTES, one env per lane); the 32 rows are contiguous in global memory, so the
// warp then streams the tile out with 128-bit stores (st.global.cs: written once, never re-read
// by this kernel) and re-zeroes the tile on the way.  Falls back to byte stores when the
// destination or the length is not 16-byte aligned (ragged last warp).
// ------------------------------------------------------------------------------------------
// LANES = lanes that share the tile (32: the whole warp; 16: a half-warp that hosts its own env, kernels_warp.cuh)
template <int LANES = kWarp>
__device__ __forceinline__ void warp_tile_zero(uint8_t *tile, int nbytes, int lane) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = lane; c < (nbytes >> 4); c += LANES) reinterpret_cast<uint4 *>(tile)[c] = z;
    for (int c = (nbytes & ~15) + lane; c < nbytes; c += LANES) tile[c] = 0;
}
template <int LANES = kWarp>
__device__ __forceinline__ void warp_tile_flush(uint8_t *gdst, uint8_t *tile, int nbytes, int lane) {
    if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)nbytes) & 15u) == 0) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int c = lane; c < (nbytes >> 4); c += LANES) {
            const uint4 v = reinterpret_cast<uint4 *>(tile)[c];
            st_stream(reinterpret_cast<uint4 *>(gdst) + c, v);
            reinterpret_cast<uint4 *>(tile)[c] = z;
        }
    } else {
        for (int c = lane; c < nbytes; c += LANES) { gdst[c] = tile[c]; tile[c] = 0; }
    }
}

// Full-tile fast path: TILE_BYTES (compile time, a multiple of 16) to a 16-byte aligned destination -> the trip
// count and every guard are compile-time.  Chunks move in batches of up to four per lane: all 128-bit LDS of a batch
// first, then the zeroing STS, then the STG.cs -- distinct registers per chunk, so no load waits for a store to
// release its source registers and the shared-memory latency is paid once per batch.
template <int TILE_BYTES, int LANES = kWarp>
__device__ __forceinline__ void warp_tile_flush_full(uint8_t *gdst, uint8_t *tile, int lane) {
    static_assert(TILE_BYTES % 16 == 0, "tile must be a whole number of 128-bit chunks");
    constexpr int kWarp = LANES;           // lanes that share the tile (shadows rlc::kWarp in this function)
    constexpr int kChunks = TILE_BYTES / 16, kBatch = 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4 *t4 = reinterpret_cast<uint4 *>(tile), *g4 = reinterpret_cast<uint4 *>(gdst);
#pragma unroll
    for (int b0 = 0; b0 < kChunks; b0 += kBatch * kWarp) {
        uint4 v[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; q++) {
            const int c0 = b0 + q * kWarp, c = c0 + lane;
            if (c0 < kChunks && (c0 + kWarp <= kChunks || c < kChunks)) v[q] = t4[c];
        }
#pragma unroll
        for (int q = 0; q < kBatch; q++) {
            const int c0 = b0 + q * kWarp, c = c0 + lane;
            if (c0 < kChunks && (c0 + kWarp <= kChunks || c < kChunks)) t4[c] = z;
        }
#pragma unroll
        for (int q = 0; q < kBatch; q++) {
            const int c0 = b0 + q * kWarp, c = c0 + lane;
            if (c0 < kChunks && (c0 + kWarp <= kChunks || c < kChunks)) st_stream(g4 + c, v[q]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tile store in two halves, so that the copy engine can take the flush off the instruction stream:
//   tile_store_begin   hands the (full, 16-byte aligned) tile to the destination
//   tile_store_end     returns when the tile may be written again; the tile is all zero then
// TMA = false (default, RLC_TMA_FLUSH = 0): begin = warp_tile_flush_full (LDS.128 + STS.128 zero + STG.128.cs per chunk and lane), end = nothing.
// TMA = true (RLC_TMA_FLUSH = 1 builds it everywhere): begin = one cp.async.bulk.global.shared::cta (SASS UBLKCP.G.S) issued by lane 0 after the writer
//   lanes have fenced their generic-proxy stores into the async proxy, L2 evict-first hint (the trajectory is written once
//   and never re-read); end = cp.async.bulk.wait_group.read 0 by lane 0, then the lanes re-zero the tile.  The transition
//   code that runs between begin and end overlaps the copy.
// Measured on B200 (profiles/r02_tma_flush.md): only the tabulated Leduc rollout gains (0.0915 -> 0.0906 ms); the
// register engines lose 0.5-12 % (fence + wait + re-zero sit on their latency-bound critical path), so only
// k_rollout_leduc_fsm asks for TMA = true.
// ------------------------------------------------------------------------------------------
#ifndef RLC_TMA_FLUSH
#define RLC_TMA_FLUSH 0
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// the pieces of a bulk store, for kernels that issue several rows under one fence / one commit group
__device__ __forceinline__ uint64_t bulk_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // every writer lane
__device__ __forceinline__ void bulk_issue(void *gdst, const void *tile, uint32_t bytes, uint64_t pol) {           // one lane
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(smem_u32(tile)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int TILE_BYTES, bool TMA = (RLC_TMA_FLUSH != 0)>
__device__ __forceinline__ void tile_store_begin(uint8_t *gdst, uint8_t *tile, int lane) {
  if constexpr (TMA) {
    static_assert(TILE_BYTES % 16 == 0, "bulk copies move whole 16-byte units");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(gdst), "r"(smem_u32(tile)), "r"(TILE_BYTES), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else {
    warp_tile_flush_full<TILE_BYTES>(gdst, tile, lane);
  }
}
template <int TILE_BYTES, bool TMA = (RLC_TMA_FLUSH != 0)>
__device__ __forceinline__ void tile_store_end(uint8_t *tile, int lane) {
  if constexpr (TMA) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    const uint4 z = make_uint4(0, 0, 0, 0);
    constexpr int kChunks = TILE_BYTES / 16;
#pragma unroll
    for (int c0 = 0; c0 < kChunks; c0 += kWarp) {
        const int c = c0 + lane;
        if (c0 + kWarp <= kChunks || c < kChunks) reinterpret_cast<uint4 *>(tile)[c] = z;
    }
  } else {
    (void)tile; (void)lane;
  }
}

// per-env header words stored in front of the game words of the state
struct EnvHeader {
    uint32_t episode;   // episodes started (0 = never reset); the running episode has index episode-1
    uint32_t t;         // env-steps taken in the running episode
    uint32_t k;         // env-steps taken since the env was created (Philox step index)
    __device__ __forceinline__ void load(const uint32_t *st, size_t n, size_t i) { episode = st[i]; t = st[n + i]; k = st[2 * n + i]; }
    __device__ __forceinline__ void store(uint32_t *st, size_t n, size_t i) const { st[i] = episode; st[n + i] = t; st[2 * n + i] = k; }
};
constexpr int kHeaderWords = 3;

}  // namespace rlc
