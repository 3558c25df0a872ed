// kernels_warp.cuh -- one-env-per-WARP kernels for the games with combinatorial move generation
// (Scout, DouDizhu).  8192 envs per GPU would be 256 warps with a thread per env (1.7 warps per SM);
// a warp per env gives 8192 warps, lets the 32 lanes test 32 candidate actions at a time (one
// __ballot_sync == one 32-bit word of the legal mask) and write the obs row with coalesced stores.
//
// The env state is replicated: every lane executes the (scalar) transition identically, so there is no
// intra-warp hand-off.  Small fields live in the registers of all 32 lanes; the 64-bit hands that are
// indexed by a run-time seat live in the warp's slice of shared memory, where every lane writes the same
// value (one LDS instead of a select chain, and 12-16 registers fewer: occupancy).  Random draws are made
// by lane 0 and broadcast.  State rows are array-of-structs uint32 [n][state_words] (one env = one
// contiguous row, read with warp-uniform loads).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace rlc {

constexpr uint32_t kFull = 0xffffffffu;

// lane 0 owns the chance source; results are broadcast so that all lanes stay in lock step
// LPE = lanes per env (32, or 16 when a half-warp hosts the env: `lane` is then the lane within the half and gm its mask)
template <class Ch, int LPE = 32>
struct WarpChance {
    static constexpr int kKind = Ch::kKind;
    Ch ch; int lane; uint32_t gm;
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(LPE == 32 ? kFull : gm, v, 0, LPE); }
    // the same from a point all 32 lanes reach together: compile-time mask (a per-lane mask costs a MATCH.ANY loop per call)
    __device__ __forceinline__ uint32_t bcast_all(uint32_t v) const { return __shfl_sync(kFull, v, 0, LPE); }
    __device__ __forceinline__ uint32_t below(uint32_t n) {
        uint32_t v = 0;
        if (lane == 0) v = ch.below(n);
        return bcast(v);
    }
    __device__ __forceinline__ void skip_fy(int a, int b) { if (lane == 0) ch.skip_fy(a, b); }
    __device__ __forceinline__ void begin_step(uint32_t k) { if (lane == 0) (void)ch.begin_step(k); }
    __device__ __forceinline__ void begin_reset(uint32_t k) { if (lane == 0) ch.begin_reset(k); }
    __device__ __forceinline__ void begin_episode(uint32_t e) { if (lane == 0) ch.begin_episode(e); }
    __device__ __forceinline__ int err() { return (int)bcast_all((uint32_t)ch.err); }
};
template <class Ch, int LPE>
__device__ __forceinline__ void wchance_open(WarpChance<Ch, LPE> &w, const KParams &p, size_t env, int lane, uint32_t gm = kFull) {
    w.lane = lane; w.gm = gm;
    ChanceIO<Ch>::open(w.ch, p, env);     // every lane builds the (cheap) handle; only lane 0 draws/commits
}
template <class Ch, int LPE>
__device__ __forceinline__ void wchance_close(WarpChance<Ch, LPE> &w, const KParams &p, size_t env) {
    if (w.lane == 0) ChanceIO<Ch>::close(w.ch, p, env);
}
template <class Ch, int LPE>
__device__ __forceinline__ uint32_t wpolicy_word(WarpChance<Ch, LPE> &w, const KParams &p, size_t env, uint32_t k) {
    uint32_t word = 0;
    if (w.lane == 0) {
        if constexpr (Ch::kKind == 0) word = w.ch.begin_step(k);
        else word = policy_word_only(p, env, k);
    }
    return w.bcast_all(word);
}

// mask row of one env: dense uint8 [A] (4 ids per 32-bit store) or bit-packed uint32 [W]
template <class G, int LANES = 32>
__device__ __forceinline__ void warp_write_mask(void *gmask, size_t row, const uint32_t *sm, int lane) {
    if constexpr (G::kMaskBitpacked) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(gmask) + row * (size_t)G::MASK_WORDS;
        for (int wi = lane; wi < G::MASK_WORDS; wi += LANES) st_stream(dst + wi, sm[wi]);
    } else {
        static_assert(G::A % 4 == 0, "dense mask rows are written as 32-bit words");
        uint32_t *dst = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(gmask) + row * (size_t)G::A);
#pragma unroll
        for (int q = lane; q < G::A / 4; q += LANES) {
            const uint32_t b = (sm[q >> 3] >> ((q & 7) * 4)) & 15u;
            st_stream(dst + q, (b * 0x00204081u) & 0x01010101u);       // bit a -> byte a
        }
    }
}

// rollout form of the row flush: begin hands the row to the destination, end makes the (zeroed) row writable again
template <class G, class ObsT>
__device__ __forceinline__ bool warp_row_store_begin(void *gobs, size_t row, ObsT *srow, int lane) {
    constexpr int kBytes = G::OBS * (int)sizeof(ObsT);
    uint8_t *dst = reinterpret_cast<uint8_t *>(gobs) + row * (size_t)kBytes;
#if RLC_TMA_FLUSH
    if constexpr (kBytes % 16 == 0) {
        if ((reinterpret_cast<uintptr_t>(gobs) & 15u) == 0) { tile_store_begin<kBytes>(dst, reinterpret_cast<uint8_t *>(srow), lane); return true; }
    }
#endif
    if constexpr (G::kRowFlushFull && kBytes % 16 == 0) {
        if ((reinterpret_cast<uintptr_t>(gobs) & 15u) == 0) { warp_tile_flush_full<kBytes>(dst, reinterpret_cast<uint8_t *>(srow), lane); return false; }
    }
    warp_tile_flush(dst, reinterpret_cast<uint8_t *>(srow), kBytes, lane);
    return false;
}
template <class G, class ObsT>
__device__ __forceinline__ void warp_row_store_end(ObsT *srow, bool pending, int lane) {
#if RLC_TMA_FLUSH
    constexpr int kBytes = G::OBS * (int)sizeof(ObsT);
    if constexpr (kBytes % 16 == 0) { if (pending) { tile_store_end<kBytes>(reinterpret_cast<uint8_t *>(srow), lane); __syncwarp(); } }
#else
    (void)srow; (void)pending; (void)lane;
#endif
}

template <class G, class ObsT>
__device__ __forceinline__ void warp_flush_row(void *gobs, size_t row, ObsT *srow, int lane) {
    constexpr int kBytes = G::OBS * (int)sizeof(ObsT);
    uint8_t *dst = reinterpret_cast<uint8_t *>(gobs) + row * (size_t)kBytes;
    if constexpr (G::kRowFlushFull && kBytes % 16 == 0) {      // compile-time batched flush (costs registers: per game)
        if ((reinterpret_cast<uintptr_t>(gobs) & 15u) == 0) { warp_tile_flush_full<kBytes>(dst, reinterpret_cast<uint8_t *>(srow), lane); return; }
    }
    warp_tile_flush(dst, reinterpret_cast<uint8_t *>(srow), kBytes, lane);
}

template <class G, class Ch, class ObsT, int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_wenv(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t env = (size_t)blockIdx.x * (BLOCK / 32) + wib;
    if (env >= p.n) return;
    constexpr int kRowBytes = (G::OBS * (int)sizeof(ObsT) + 15) & ~15;
    constexpr int kWarpBytes = kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15) + G::kScratchBytes;
    uint8_t *base = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)wib * kWarpBytes;
    ObsT *srow = reinterpret_cast<ObsT *>(base);
    uint32_t *smask = reinterpret_cast<uint32_t *>(base + kRowBytes);
    uint8_t *scratch = base + kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15);
    warp_tile_zero(base, kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15), lane);      // obs row and mask words
    __syncwarp();

    uint32_t *row = p.state + env * (size_t)(kHeaderWords + G::GAME_WORDS);
    EnvHeader h; h.episode = row[0]; h.t = row[1]; h.k = row[2];
    G g; g.bind(p, scratch); g.load(row + kHeaderWords, lane);
    WarpChance<Ch> ch; wchance_open(ch, p, env, lane);
    int err = 0; bool done = false;
    float pay[G::P];
#pragma unroll
    for (int k = 0; k < G::P; k++) pay[k] = 0.f;

    if constexpr (MODE == kModeReset) {
        if (!p.reset_mask || p.reset_mask[env]) {
            ch.begin_reset(h.k);
            h.episode++; h.t = 0;
            g.reset(ch, scratch, lane);
        }
    } else if constexpr (MODE == kModeStep) {
        const int a = p.actions[env];
        if (a >= 0 && h.episode != 0 && !g.over()) {
            ch.begin_step(h.k);
            g.legal(smask, scratch, lane);                        // the step validates against the current legal set
            __syncwarp();
            g.step(a, ch, smask, scratch, lane, err);
            h.t++; h.k++;
            if (g.over()) {
                done = true;
                g.payoffs(pay);
                if ((p.flags & RLC_TERMINAL_OBS) && p.terminal_obs) {
                    g.legal(smask, scratch, lane);
                    __syncwarp();
                    for (int s = 0; s < G::P; s++) {
                        g.encode_obs(s, false, srow, scratch, lane);
                        __syncwarp();
                        warp_flush_row<G, ObsT>(p.terminal_obs, env * G::P + s, srow, lane);
                        __syncwarp();
                    }
                }
                if (p.flags & RLC_AUTO_RESET) { h.episode++; h.t = 0; g.reset(ch, scratch, lane); }
            }
        } else if (g.over()) { done = true; g.payoffs(pay); }
    } else {
        if (h.episode != 0 && g.over()) { done = true; g.payoffs(pay); }
    }
    __syncwarp();
    g.legal(smask, scratch, lane);
    __syncwarp();
    const int seat = (MODE == kModeObserve && p.seat) ? p.seat[env] : g.player();
    if (p.obs) {
        g.encode_obs(seat, MODE != kModeObserve && h.t == 0, srow, scratch, lane);
        __syncwarp();
        warp_flush_row<G, ObsT>(p.obs, env, srow, lane);
    }
    if (p.mask) warp_write_mask<G>(p.mask, env, smask, lane);
    if (p.order) {
        __syncwarp();
        g.legal_order(smask, p.order + env * (size_t)p.order_stride, p.order_stride, lane);
    }
    if (lane == 0) {
        if (p.cur_player) p.cur_player[env] = g.player();
        if (p.done) p.done[env] = done ? 1 : 0;
        if (p.payoffs) {
#pragma unroll
            for (int k = 0; k < G::P; k++) p.payoffs[env * G::P + k] = pay[k];
        }
    }
    if constexpr (MODE != kModeObserve) {
        wchance_close(ch, p, env);
        if (lane == 0) { row[0] = h.episode; row[1] = h.t; row[2] = h.k; }
        g.store(row + kHeaderWords, lane);
    }
    err |= ch.err();
    if (err && p.err && lane == 0) p.err[env] |= err;
}

// EXT = the rlc_trajectory extensions are in use (recorded actions instead of the random policy, terminal-state pool);
// the plain instantiation is the throughput kernel and carries none of that code.
// BULK (throughput kernel only; RLC_WROLLOUT_BULK picks it at launch, default per game from the B200 measurements in
// profiles/): bit 0 = the bit-packed mask row leaves shared memory as ONE cp.async.bulk copy (UBLKCP) instead of 27 LDS + STG
// per lane, bit 1 = the obs row too.  The warp-per-env kernels are issue-bound, so every instruction the copy engine takes
// over counts; the mask row is only rewritten by the next legal(), ~300 instructions later, so the wait is free.
template <class G, class Ch, class ObsT, int BLOCK, bool EXT, int BULK = 0>
__global__ void __launch_bounds__(BLOCK, G::kMinBlocks) k_wrollout(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t env = (size_t)blockIdx.x * (BLOCK / 32) + wib;
    if (env >= p.n) return;
    constexpr int kRowBytes = (G::OBS * (int)sizeof(ObsT) + 15) & ~15;
    constexpr int kWarpBytes = kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15) + G::kScratchBytes;
    uint8_t *base = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)wib * kWarpBytes;
    ObsT *srow = reinterpret_cast<ObsT *>(base);
    uint32_t *smask = reinterpret_cast<uint32_t *>(base + kRowBytes);
    uint8_t *scratch = base + kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15);
    warp_tile_zero(base, kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15), lane);      // obs row and mask words
    __syncwarp();

    uint32_t *row = p.state + env * (size_t)(kHeaderWords + G::GAME_WORDS);
    EnvHeader h; h.episode = row[0]; h.t = row[1]; h.k = row[2];
    G g; g.bind(p, scratch); g.load(row + kHeaderWords, lane);
    WarpChance<Ch> ch; wchance_open(ch, p, env, lane);
    int err = 0;
    if (h.episode == 0 || g.over()) { ch.begin_reset(h.k); h.episode++; h.t = 0; g.reset(ch, scratch, lane); }
    __syncwarp();
    int cnt = g.legal(smask, scratch, lane);
    __syncwarp();
    size_t rowi = env;
    constexpr int kObsBytes = G::OBS * (int)sizeof(ObsT), kMaskBytes = G::MASK_WORDS * 4;
    constexpr bool kBulkMask = (BULK & 1) != 0 && G::kMaskBitpacked && kMaskBytes % 16 == 0;
    constexpr bool kBulkObs = (BULK & 2) != 0 && kBulkMask && kObsBytes % 16 == 0;          // the obs row only together with the mask row
    uint64_t bulk_pol = 0;
    if constexpr (kBulkMask) bulk_pol = bulk_evict_first_policy();
    for (int t = 0; t < p.T; t++, rowi += p.n) {
        bool row_pending = false;
        if constexpr (kBulkObs && kBulkMask) {                    // launcher: both streams present and aligned
            g.encode_obs(g.player(), h.t == 0, srow, scratch, lane);
            bulk_fence();                                         // one fence, one barrier, one commit group for both rows
            __syncwarp();
            if (lane == 0) {
                bulk_issue(reinterpret_cast<uint8_t *>(p.t_obs) + rowi * (size_t)kObsBytes, srow, kObsBytes, bulk_pol);
                bulk_issue(reinterpret_cast<uint8_t *>(p.t_mask) + rowi * (size_t)kMaskBytes, smask, kMaskBytes, bulk_pol);
                bulk_commit();
            }
        } else {
            if (p.t_obs) {
                g.encode_obs(g.player(), h.t == 0, srow, scratch, lane);
                __syncwarp();
                row_pending = warp_row_store_begin<G, ObsT>(p.t_obs, rowi, srow, lane);
            }
            if (p.t_mask) {
                if constexpr (kBulkMask) {
                    tile_store_begin<kMaskBytes, true>(reinterpret_cast<uint8_t *>(p.t_mask) + rowi * (size_t)kMaskBytes, reinterpret_cast<uint8_t *>(smask), lane);
                } else warp_write_mask<G>(p.t_mask, rowi, smask, lane);
            }
        }
        const uint32_t word = wpolicy_word(ch, p, env, h.k);
        const int k = (int)__umulhi(word, (uint32_t)cnt);
        int a;
        if (EXT && p.t_forced) a = p.t_forced[rowi];              // replay of a recorded action sequence (< 0: env idles)
        else a = g.pick(smask, scratch, k, lane);
        const int pl = g.player();
        __syncwarp();
        bool over = false;
        float pay[G::P];
#pragma unroll
        for (int q = 0; q < G::P; q++) pay[q] = 0.f;
        if (!EXT || a >= 0) {
            g.step(a, ch, smask, scratch, lane, err);
            h.t++; h.k++;
            over = g.over();
        }
        if constexpr (EXT) { warp_row_store_end<G, ObsT>(srow, row_pending, lane); row_pending = false; }   // srow is reused below
        if (EXT && p.tm_row) {                                    // per-seat terminal states into the pool (env.py:161-164)
            int r = -1;
            if (over) {
                if (lane == 0) r = atomicAdd(p.tm_count, 1);
                r = __shfl_sync(kFull, r, 0);
                if (r >= p.tm_cap) { r = -1; err |= kErrTerminalPoolFull; }
            }
            if (lane == 0) p.tm_row[rowi] = r;
            if (r >= 0) {
                __syncwarp();
                g.legal(smask, scratch, lane);
                __syncwarp();
                if (p.tm_obs) {
                    for (int s = 0; s < G::P; s++) {
                        g.encode_obs(s, false, srow, scratch, lane);
                        __syncwarp();
                        warp_flush_row<G, ObsT>(p.tm_obs, (size_t)r * G::P + s, srow, lane);
                        __syncwarp();
                    }
                }
                if (p.tm_mask) warp_write_mask<G>(p.tm_mask, (size_t)r, smask, lane);
                __syncwarp();
            }
        }
        if (over) { g.payoffs(pay); h.episode++; h.t = 0; g.reset(ch, scratch, lane); }
        if constexpr (kBulkObs && kBulkMask) tile_store_end<kObsBytes, true>(reinterpret_cast<uint8_t *>(srow), lane);   // wait + re-zero the obs row
        else if constexpr (kBulkMask) { if (p.t_mask) { if (lane == 0) bulk_wait_read(); } }
        __syncwarp();
        if (!EXT || a >= 0) cnt = g.legal(smask, scratch, lane);  // legal set of the state the next iteration emits
        __syncwarp();
        warp_row_store_end<G, ObsT>(srow, row_pending, lane);
        if (lane == 0) {
            if (p.t_player) st_stream(p.t_player + rowi, pl);
            if (p.t_action) st_stream(p.t_action + rowi, a);
            if (p.t_done) p.t_done[rowi] = over ? 1 : 0;
        }
        if (p.t_payoffs && lane < G::P) {
            float v = pay[0];
#pragma unroll
            for (int q = 1; q < G::P; q++) v = lane == q ? pay[q] : v;
            p.t_payoffs[rowi * G::P + lane] = v;
        }
    }
    wchance_close(ch, p, env);
    if (lane == 0) { row[0] = h.episode; row[1] = h.t; row[2] = h.k; }
    g.store(row + kHeaderWords, lane);
    err |= ch.err();
    if (err && p.err && lane == 0) p.err[env] |= err;
}

// Throughput rollout with LPE < 32 lanes per env: a warp hosts 32 / LPE envs, one per lane group.  The transition is scalar
// code that every lane of an env replicates, so a warp-instruction of it now serves 32 / LPE envs instead of one; the
// lane-parallel parts (legal-set ballots, obs slots, row flush) run on the group's lanes.  Groups diverge where their envs do
// (play vs scout, an episode ending), so every barrier and collective names the group's own lane mask.  Philox mode, every
// trajectory stream optional, no ABI-2 extensions (those run on k_wrollout).
template <class G, class Ch, class ObsT, int BLOCK, int LPE>
__global__ void __launch_bounds__(BLOCK, (LPE == 16 && BLOCK == 128) ? G::kMinBlocks : (BLOCK <= 64 ? 8 : (BLOCK <= 128 ? 4 : (BLOCK <= 256 ? 2 : 1)))) k_wrollout_multi(const KParams p) {
    static_assert((LPE == 16 || LPE == 8) && G::kLanes == LPE, "half- or quarter-warp groups");
    extern __shared__ uint4 smem_raw[];
    constexpr int kGroups = 32 / LPE;
    const int lane_abs = threadIdx.x & 31, sub = lane_abs / LPE, lane = lane_abs % LPE;
    const int slot = (threadIdx.x >> 5) * kGroups + sub;
    const size_t env0 = (size_t)blockIdx.x * (BLOCK / LPE) + slot;
    if (env0 - sub >= p.n) return;                                // whole warp past the batch
    // the collectives of the transition name all 32 lanes, so a group past the end of a ragged batch stays: it replays the
    // last env in its own shared memory and stores nothing
    const bool live = env0 < p.n;
    const size_t env = live ? env0 : p.n - 1;
    const int gsh = sub * LPE;
    const uint32_t gm = ((1u << LPE) - 1u) << gsh;
    constexpr int kRowBytes = (G::OBS * (int)sizeof(ObsT) + 15) & ~15;
    constexpr int kEnvBytes = kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15) + G::kScratchBytes;
    uint8_t *base = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)slot * kEnvBytes;
    ObsT *srow = reinterpret_cast<ObsT *>(base);
    uint32_t *smask = reinterpret_cast<uint32_t *>(base + kRowBytes);
    uint8_t *scratch = base + kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15);
    warp_tile_zero<LPE>(base, kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15), lane);
    __syncwarp();

    uint32_t *row = p.state + env * (size_t)(kHeaderWords + G::GAME_WORDS);
    EnvHeader h; h.episode = row[0]; h.t = row[1]; h.k = row[2];
    G g; g.bind(p, scratch); g.bind_group(gm, gsh); g.load(row + kHeaderWords, lane);
    WarpChance<Ch, LPE> ch; wchance_open(ch, p, env, lane, gm);
    int err = 0;
    if (h.episode == 0 || g.over()) { ch.begin_reset(h.k); h.episode++; h.t = 0; g.reset(ch, scratch, lane); }
    __syncwarp();
    int cnt = g.legal(smask, scratch, lane);
    __syncwarp();
    size_t rowi = env;
    constexpr int kObsBytes = G::OBS * (int)sizeof(ObsT);
    const bool obs_fast = kObsBytes % 16 == 0 && (reinterpret_cast<uintptr_t>(p.t_obs) & 15u) == 0;
    for (int t = 0; t < p.T; t++, rowi += p.n) {
        if (p.t_obs) {
            g.encode_obs(g.player(), h.t == 0, srow, scratch, lane);
            __syncwarp();
            uint8_t *dst = reinterpret_cast<uint8_t *>(p.t_obs) + rowi * (size_t)kObsBytes;
            if (!live) warp_tile_zero<LPE>(reinterpret_cast<uint8_t *>(srow), kRowBytes, lane);
            else if constexpr (kObsBytes % 16 == 0) {
                if (obs_fast) warp_tile_flush_full<kObsBytes, LPE>(dst, reinterpret_cast<uint8_t *>(srow), lane);
                else warp_tile_flush<LPE>(dst, reinterpret_cast<uint8_t *>(srow), kObsBytes, lane);
            } else warp_tile_flush<LPE>(dst, reinterpret_cast<uint8_t *>(srow), kObsBytes, lane);
        }
        if (p.t_mask && live) warp_write_mask<G, LPE>(p.t_mask, rowi, smask, lane);
        const uint32_t word = wpolicy_word(ch, p, env, h.k);
        const int k = (int)__umulhi(word, (uint32_t)cnt);
        const int a = g.pick(smask, scratch, k, lane);
        const int pl = g.player();
        __syncwarp();
        float pay[G::P];
#pragma unroll
        for (int q = 0; q < G::P; q++) pay[q] = 0.f;
        g.step(a, ch, smask, scratch, lane, err);
        h.t++; h.k++;
        const bool over = g.over();
        if (over) { g.payoffs(pay); h.episode++; h.t = 0; g.reset(ch, scratch, lane); }
        __syncwarp();
        cnt = g.legal(smask, scratch, lane);                      // legal set of the state the next iteration emits
        __syncwarp();
        if (lane == 0 && live) {
            if (p.t_player) st_stream(p.t_player + rowi, pl);
            if (p.t_action) st_stream(p.t_action + rowi, a);
            if (p.t_done) p.t_done[rowi] = over ? 1 : 0;
        }
        if (p.t_payoffs && lane < G::P && live) {
            float v = pay[0];
#pragma unroll
            for (int q = 1; q < G::P; q++) v = lane == q ? pay[q] : v;
            p.t_payoffs[rowi * G::P + lane] = v;
        }
    }
    err |= ch.err();
    if (!live) return;
    wchance_close(ch, p, env);
    if (lane == 0) { row[0] = h.episode; row[1] = h.t; row[2] = h.k; }
    g.store(row + kHeaderWords, lane);
    if (err && p.err && lane == 0) p.err[env] |= err;
}

// lanes per env of the fused throughput rollout, if the game asks for fewer than a warp (G::kRolloutLanes)
template <class G, class = void> struct RolloutLanes { static constexpr int value = 32; };
template <class G> struct RolloutLanes<G, std::void_t<decltype(G::kRolloutLanes)>> { static constexpr int value = G::kRolloutLanes; };

template <class G, class Ch, class ObsT>
cudaError_t launch_wop(int op, const KParams &p, cudaStream_t stream) {
    constexpr int BLOCK = 128;
    const unsigned grid = (unsigned)((p.n + BLOCK / 32 - 1) / (BLOCK / 32));
    constexpr int kRowBytes = (G::OBS * (int)sizeof(ObsT) + 15) & ~15;
    const size_t smem = (size_t)(BLOCK / 32) * (kRowBytes + ((G::MASK_WORDS * 4 + 15) & ~15) + G::kScratchBytes);
    cudaError_t e = cudaSuccess;
#define RLC_WLAUNCH(KERNEL)                                                                          \
    do {                                                                                             \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) { KERNEL<<<grid, BLOCK, smem, stream>>>(p); e = cudaGetLastError(); }  \
    } while (0)
    switch (op) {
    case kOpReset: RLC_WLAUNCH((k_wenv<G, Ch, ObsT, kModeReset, BLOCK>)); break;
    case kOpStep: RLC_WLAUNCH((k_wenv<G, Ch, ObsT, kModeStep, BLOCK>)); break;
    case kOpObserve: RLC_WLAUNCH((k_wenv<G, Ch, ObsT, kModeObserve, BLOCK>)); break;
    case kOpRollout:
        if (p.t_forced || p.tm_row) RLC_WLAUNCH((k_wrollout<G, Ch, ObsT, BLOCK, true>));
        else {
            if constexpr (Ch::kKind == 0 && RolloutLanes<G>::value < 32) {          // several envs per warp (throughput mode)
                const char *lv = getenv("RLC_WROLLOUT_LPE");
                const int lpe = lv ? atoi(lv) : RolloutLanes<G>::value;
                if (lpe == 16 || lpe == 8) {
                    auto go = [&](auto kk, int block, int lanes) {
                        const unsigned mgrid = (unsigned)((p.n + block / lanes - 1) / (block / lanes));
                        const size_t msmem = (smem / (BLOCK / 32)) * (size_t)(block / lanes);
                        if (msmem > 48 * 1024) e = cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
                        if (e == cudaSuccess) { kk<<<mgrid, block, msmem, stream>>>(p); e = cudaGetLastError(); }
                    };
                    // a quarter-warp per env: 64-thread blocks (8 envs, 23 KB of shared memory for Scout) keep the SMs evenly filled
                    // block size at 8 lanes per env: 64 threads (8 envs) spread any batch evenly; when the per-SM share of the batch is
                    // 33..56 envs, ONE 448-thread block per SM (14 warps) does the same with every SM equally loaded (Scout, 8 192 envs:
                    // 0.617 -> 0.605 ms; 128 / 256 / 512 threads: 0.643 / 0.617 / 0.608).  RLC_WROLLOUT_BLOCK overrides.
                    static int sms = 0;
                    if (sms == 0) {
                        int dev = 0, v = 148;
                        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
                        sms = v > 0 ? v : 148;
                    }
                    const size_t share = (p.n + (size_t)sms - 1) / (size_t)sms;
                    const char *bv = getenv("RLC_WROLLOUT_BLOCK");
                    const int blk = bv ? atoi(bv) : ((share > 32 && share <= 56) ? 448 : 64);
                    if (lpe == 8 && blk == 128) go(k_wrollout_multi<typename G::template WithLanes<8>, Ch, ObsT, 128, 8>, 128, 8);
                    else if (lpe == 8 && blk == 256) go(k_wrollout_multi<typename G::template WithLanes<8>, Ch, ObsT, 256, 8>, 256, 8);
                    else if (lpe == 8 && blk == 448) go(k_wrollout_multi<typename G::template WithLanes<8>, Ch, ObsT, 448, 8>, 448, 8);
                    else if (lpe == 8 && blk == 512) go(k_wrollout_multi<typename G::template WithLanes<8>, Ch, ObsT, 512, 8>, 512, 8);
                    else if (lpe == 8) go(k_wrollout_multi<typename G::template WithLanes<8>, Ch, ObsT, 64, 8>, 64, 8);
                    else go(k_wrollout_multi<typename G::template WithLanes<16>, Ch, ObsT, BLOCK, 16>, BLOCK, 16);
                    break;
                }
            }
            int bulk = 0;
            if constexpr (Ch::kKind == 0 && G::kMaskBulk) {
                constexpr int kObsBytes = G::OBS * (int)sizeof(ObsT), kMaskBytes = G::MASK_WORDS * 4;
                const char *envv = getenv("RLC_WROLLOUT_BULK");
                bulk = envv ? atoi(envv) : G::kBulkDefault;
                if ((reinterpret_cast<uintptr_t>(p.t_mask) & 15u) || kMaskBytes % 16) bulk &= ~1;
                if ((reinterpret_cast<uintptr_t>(p.t_obs) & 15u) || kObsBytes % 16) bulk &= ~2;
                if (!p.t_mask) bulk &= ~1;
                if (!p.t_obs) bulk &= ~2;
                if (bulk == 1) { RLC_WLAUNCH((k_wrollout<G, Ch, ObsT, BLOCK, false, 1>)); break; }
                if (bulk == 3) { RLC_WLAUNCH((k_wrollout<G, Ch, ObsT, BLOCK, false, 3>)); break; }
            }
            RLC_WLAUNCH((k_wrollout<G, Ch, ObsT, BLOCK, false>));
        }
        break;
    default: e = cudaErrorInvalidValue;
    }
#undef RLC_WLAUNCH
    return e;
}

template <class G, class ObsT>
cudaError_t dispatch_wgame(int op, int chance, const KParams &p, cudaStream_t stream) {
    if (chance == RLC_CHANCE_PHILOX) return launch_wop<G, ChancePhilox, ObsT>(op, p, stream);
    if (chance == RLC_CHANCE_TAPE) return launch_wop<G, ChanceTape, ObsT>(op, p, stream);
    if (chance == RLC_CHANCE_MT19937) return launch_wop<G, ChanceMt, ObsT>(op, p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace rlc
