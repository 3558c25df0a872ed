// kernels.cuh -- one-env-per-thread kernels, templated on <Game, Chance source, obs dtype>.
//   k_env<RESET|STEP|OBSERVE>  Env.reset / Env.step / Env.get_state over n envs (env.py:52-86,188-197)
//   k_rollout                  the Env.run loop with RandomAgents fused on device (env.py:120-169)
// State is SoA uint32 [words][n] (coalesced word loads); obs rows are staged per warp in shared
// memory and streamed out with 128-bit stores (common.cuh warp_tile_flush).
#pragma once
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "../../include/rlcard_b200.h"

namespace rlc {

struct KParams {
    uint32_t *state; size_t n;
    uint64_t seed; uint32_t env_id_base;
    const uint8_t *tape; int tape_stride; int32_t *tape_pos; uint32_t *mt;
    void *obs; void *mask; int32_t *cur_player; uint8_t *done; float *payoffs; void *terminal_obs; int32_t *err;
    const int32_t *actions; const uint8_t *reset_mask; const int32_t *seat; int flags;
    void *t_obs; void *t_mask; int32_t *t_action; int32_t *t_player; uint8_t *t_done; float *t_payoffs; int T;
    // fused-rollout extensions (rlc_trajectory, ABI 2): recorded action ids to apply instead of the random policy, and the
    // pool of per-seat terminal states (env.py:161-164) with the row index per trajectory cell
    const int32_t *t_forced; void *tm_obs; void *tm_mask; int32_t *tm_row; int32_t *tm_count; int tm_cap;
    int32_t *order; int order_stride;            // legal ids in the reference's insertion order (rlc_buffers.legal_order)
    const void *tables; const void *tab[6];      // per-game constant tables (device pointers), see Game::bind
};

enum { kModeReset = 0, kModeStep = 1, kModeObserve = 2 };
enum { kFlagNoFsm = 0x100 };         // internal launch flag: extensions present -> generic kernels only
enum { kErrTerminalPoolFull = 8 };

// ---- chance source plumbing -------------------------------------------------------------
template <class Ch> struct ChanceIO;
template <> struct ChanceIO<ChancePhilox> {
    static __device__ __forceinline__ void open(ChancePhilox &c, const KParams &p, size_t i) {
        c.init(p.seed, p.env_id_base + (uint32_t)i);
    }
    static __device__ __forceinline__ void close(ChancePhilox &, const KParams &, size_t) {}
};
// base word W_k of step k for the random policy when the chance draws come from a tape / MT19937
__device__ __forceinline__ uint32_t policy_word_only(const KParams &p, size_t i, uint32_t k) {
    uint32_t o0, o1, o2, o3;
    philox4x32_10(k >> 2, 0u, p.env_id_base + (uint32_t)i, (uint32_t)kDomBase, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), o0, o1, o2, o3);
    return sel4(o0, o1, o2, o3, k & 3u);
}
template <> struct ChanceIO<ChanceTape> {
    static __device__ __forceinline__ void open(ChanceTape &c, const KParams &p, size_t i) {
        c.tape = p.tape + i * (size_t)p.tape_stride; c.len = p.tape_stride; c.pos = p.tape_pos[i]; c.err = 0;
    }
    static __device__ __forceinline__ void close(ChanceTape &c, const KParams &p, size_t i) { p.tape_pos[i] = c.pos; }
};
template <> struct ChanceIO<ChanceMt> {
    static __device__ __forceinline__ void open(ChanceMt &c, const KParams &p, size_t i) {
        c.mt = p.mt + i; c.n = p.n; c.mti = (int)p.mt[(size_t)624 * p.n + i]; c.err = 0;
    }
    static __device__ __forceinline__ void close(ChanceMt &c, const KParams &p, size_t i) {
        p.mt[(size_t)624 * p.n + i] = (uint32_t)c.mti;
    }
};

// games whose packed state depends on the chance kind (Blackjack: deck mask vs ordered deck) expose
// load_k/store_k<kind>; the others plain load/store
template <class G, class Ch>
__device__ __forceinline__ void game_load(G &g, const uint32_t *st, size_t n, size_t i) {
    if constexpr (G::kChanceAwareState) g.template load_k<Ch::kKind>(st, n, i);
    else g.load(st, n, i);
}
template <class G, class Ch>
__device__ __forceinline__ void game_store(const G &g, uint32_t *st, size_t n, size_t i) {
    if constexpr (G::kChanceAwareState) g.template store_k<Ch::kKind>(st, n, i);
    else g.store(st, n, i);
}

// deal the next episode; the chance draws continue the sequence of the current step
template <class G, class Ch>
__device__ __forceinline__ void new_episode(G &g, Ch &ch, EnvHeader &h) {
    h.episode++; h.t = 0;
    ch.begin_episode(h.episode);          // deal words are keyed by the episode ordinal (games with kEpisodeDeal)
    g.reset(ch);
}

// dense uint8 mask row [A] of one env; rows of a warp are contiguous
template <class G>
__device__ __forceinline__ void write_mask_row(uint8_t *gmask, size_t env, const uint32_t (&m)[G::MASK_WORDS]) {
    if constexpr (G::A == 4) {
        const uint32_t w = (m[0] & 1u) | ((m[0] & 2u) << 7) | ((m[0] & 4u) << 14) | ((m[0] & 8u) << 21);
        st_stream(reinterpret_cast<uint32_t *>(gmask) + env, w);
    } else if constexpr (G::A == 2) {
        reinterpret_cast<uint16_t *>(gmask)[env] = (uint16_t)((m[0] & 1u) | ((m[0] & 2u) << 7));
    } else {
        uint8_t *rowp = gmask + env * (size_t)G::A;
        for (int a = 0; a < G::A; a++) rowp[a] = (m[a >> 5] >> (a & 31)) & 1u;
    }
}

// uniform-random legal action: k-th legal id in ascending order, k = mulhi(policy word, #legal)
template <class G>
__device__ __forceinline__ int pick_action(const uint32_t (&m)[G::MASK_WORDS], uint32_t word, int &cnt) {
    cnt = popc_words<G::MASK_WORDS>(m);
    const int k = (int)__umulhi(word, (uint32_t)cnt);
    if constexpr (G::A <= 5) {
        uint32_t mm = m[0];
        if (k > 0) mm &= mm - 1;
        if (k > 1) mm &= mm - 1;
        if (k > 2) mm &= mm - 1;
        if (G::A > 4 && k > 3) mm &= mm - 1;
        return __ffs(mm) - 1;
    } else {
        return kth_set_bit<G::MASK_WORDS>(m, k);
    }
}

// legal ids in the insertion order of the reference's state['legal_actions'] (see rlc_buffers.legal_order); games without
// an order of their own list ascending
template <class G, class = void> struct HasLegalOrder { static constexpr bool value = false; };
template <class G> struct HasLegalOrder<G, std::void_t<decltype(G::kHasLegalOrder)>> { static constexpr bool value = G::kHasLegalOrder; };
template <class G>
__device__ void write_legal_order(const G &g, const uint32_t (&m)[G::MASK_WORDS], int32_t *out, int stride) {
    int n = 0;
    if constexpr (HasLegalOrder<G>::value) n = g.legal_order(m, out, stride);
    else {
        for (int a = 0; a < G::A; a++)
            if ((m[a >> 5] >> (a & 31)) & 1u) { if (n < stride) out[n] = a; n++; }
    }
    for (int k = n; k < stride; k++) out[k] = -1;
}

template <class G, class Ch, class ObsT, int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_env(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int kRowBytes = G::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = BLOCK * kRowBytes;
    if constexpr (G::kSharedBytes > 0) {
        G::fill_shared(reinterpret_cast<uint8_t *>(smem_raw) + kTileBytes, threadIdx.x, BLOCK);
        __syncthreads();
    }
    const size_t warp_env0 = ((size_t)blockIdx.x * (BLOCK / 32) + wib) * 32;
    if (warp_env0 >= p.n) return;
    const size_t i = warp_env0 + lane;
    const bool valid = i < p.n;
    const int nvalid = (int)min((size_t)32, p.n - warp_env0);
    ObsT *tile = reinterpret_cast<ObsT *>(smem_raw) + (size_t)wib * 32 * G::OBS;
    ObsT *row = tile + lane * G::OBS;
    warp_tile_zero(reinterpret_cast<uint8_t *>(tile), 32 * kRowBytes, lane);
    __syncwarp();

    G g; EnvHeader h; Ch ch; int err = 0;
    g.bind_shared(reinterpret_cast<const uint8_t *>(smem_raw) + kTileBytes);
    if (valid) {
        h.load(p.state, p.n, i);
        game_load<G, Ch>(g, p.state + kHeaderWords * p.n, p.n, i);
        ChanceIO<Ch>::open(ch, p, i);
        bool done = false;
        float pay[G::P];
#pragma unroll
        for (int k = 0; k < G::P; k++) pay[k] = 0.f;
        if constexpr (MODE == kModeReset) {
            if (!p.reset_mask || p.reset_mask[i]) {
                ch.begin_reset(h.k);                  // a reset outside a step: its own Philox domain
                new_episode(g, ch, h);
            }
        } else if constexpr (MODE == kModeStep) {
            const int a = p.actions[i];
            if (a >= 0 && h.episode != 0 && !g.over()) {
                const uint32_t word = ch.begin_step(h.k);
                if constexpr (G::kUsesChain) {            // same draw derivation as the fused rollout
                    uint32_t m0[G::MASK_WORDS];
                    g.legal(m0);
                    ch.seed_chain(word, (uint32_t)popc_words<G::MASK_WORDS>(m0));
                }
                g.step(a, ch, err);
                h.t++; h.k++;
                if (g.over()) {
                    done = true;
                    g.payoffs(pay);
                    if ((p.flags & RLC_TERMINAL_OBS) && p.terminal_obs) {
                        ObsT *dst = reinterpret_cast<ObsT *>(p.terminal_obs) + i * (size_t)(G::P * G::OBS);
                        for (int s = 0; s < G::P; s++) {
                            g.encode_obs(s, false, row);
                            for (int k = 0; k < G::OBS; k++) { dst[s * G::OBS + k] = row[k]; row[k] = (ObsT)0; }
                        }
                    }
                    if (p.flags & RLC_AUTO_RESET) new_episode(g, ch, h);
                }
            } else if (g.over()) { done = true; g.payoffs(pay); }
        } else {
            if (h.episode != 0 && g.over()) { done = true; g.payoffs(pay); }
        }
        const int seat = (MODE == kModeObserve && p.seat) ? p.seat[i] : g.player();
        if (p.obs) g.encode_obs(seat, MODE != kModeObserve && h.t == 0, row);
        uint32_t m[G::MASK_WORDS];
        g.legal(m);
        if (p.mask) write_mask_row<G>(reinterpret_cast<uint8_t *>(p.mask), i, m);
        if (p.order) write_legal_order<G>(g, m, p.order + i * (size_t)p.order_stride, p.order_stride);
        if (p.cur_player) p.cur_player[i] = g.player();
        if (p.done) p.done[i] = done ? 1 : 0;
        if (p.payoffs) {
#pragma unroll
            for (int k = 0; k < G::P; k++) p.payoffs[i * G::P + k] = pay[k];
        }
        if constexpr (MODE != kModeObserve) {
            ChanceIO<Ch>::close(ch, p, i);
            h.store(p.state, p.n, i);
            game_store<G, Ch>(g, p.state + kHeaderWords * p.n, p.n, i);
        }
        err |= ch.err;
        if (err && p.err) p.err[i] |= err;
    }
    __syncwarp();
    if (p.obs)
        warp_tile_flush(reinterpret_cast<uint8_t *>(p.obs) + warp_env0 * (size_t)kRowBytes,
                        reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
}

// one env's dense mask row into the warp's mask tile (rows of a warp are contiguous in global memory)
template <class G, int EPW>
__device__ __forceinline__ void stage_mask_row(uint8_t *mtile, int lane, const uint32_t (&m)[G::MASK_WORDS]) {
    if constexpr (G::MASK_WORDS == 2 && G::A <= 61 && (EPW * G::A) % 4 == 0) {
        // 4 mask bits -> 4 bytes per 32-bit store instead of one byte store per action: the row starts at byte
        // lane * A, o = (lane * A) & 3 bytes past a word boundary, so the bits are shifted up by o and expanded
        // on the word grid; the first and the last word are shared with the neighbouring rows and OR-ed in
        // (the tile is zero between steps), the words in between are this lane's alone
        const int o = (lane * G::A) & 3;
        const uint64_t mm = ((uint64_t)m[0] | ((uint64_t)m[1] << 32)) << o;
        const uint32_t lo = (uint32_t)mm, hi = (uint32_t)(mm >> 32);
        uint32_t *w = reinterpret_cast<uint32_t *>(mtile + lane * G::A - o);
        constexpr int kW = (G::A + 3 + 3) / 4;
#pragma unroll
        for (int j = 0; j < kW; j++) {
            const uint32_t nibble = ((j < 8 ? lo : hi) >> (4 * (j & 7))) & 15u;
            const uint32_t v = (nibble * 0x00204081u) & 0x01010101u;          // bit a -> byte a
            if (j == 0 || j == kW - 1) atomicOr(w + j, v);
            else w[j] = v;
        }
    } else {
        uint8_t *mr = mtile + lane * G::A;
#pragma unroll
        for (int a = 0; a < G::A; a++) mr[a] = (m[a >> 5] >> (a & 31)) & 1u;
    }
}

// the same by lane pairs (16 envs per warp): lane L stages the first eight words of env L's row, lane L + 16 the rest
template <class G>
__device__ __forceinline__ void stage_mask_row_pair(uint8_t *mtile, int lane, bool valid, const uint32_t (&m)[G::MASK_WORDS]) {
    static_assert(G::MASK_WORDS == 2 && G::A <= 61 && (16 * G::A) % 4 == 0, "61-id rows on the word grid");
    const int src = lane & 15;
    const uint32_t m0 = __shfl_sync(0xffffffffu, m[0], src), m1 = __shfl_sync(0xffffffffu, m[1], src);
    if (!__shfl_sync(0xffffffffu, (int)valid, src)) return;
    const int o = (src * G::A) & 3;
    const uint64_t mm = ((uint64_t)m0 | ((uint64_t)m1 << 32)) << o;
    const bool upper = lane >= 16;
    const uint32_t half = upper ? (uint32_t)(mm >> 32) : (uint32_t)mm;
    uint32_t *w = reinterpret_cast<uint32_t *>(mtile + src * G::A - o) + (upper ? 8 : 0);
    constexpr int kW = (G::A + 3 + 3) / 4;                    // 16 words
    static_assert(kW == 16, "two lanes, eight words each");
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t v = (((half >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;     // bit a -> byte a
        if (j == 0 || j == 7) atomicOr(w + j, v);             // words 0 / 15 of the row are shared with its neighbours; 7 / 8 are not
        else w[j] = v;                                        // (an OR there is harmless: the tile is zero between steps)
    }
}
template <class G, class = void> struct HasPairEmit { static constexpr bool value = false; };
template <class G> struct HasPairEmit<G, std::void_t<decltype(G::kPairEmit)>> { static constexpr bool value = G::kPairEmit; };

// Fused random rollout: T env-steps per env, state in registers for the whole launch; per step the
// warp emits one coalesced obs tile plus mask/action/player/done/payoff rows of the trajectory.
// ALL = every trajectory pointer is present and the obs rows of a full warp are 16-byte aligned (the
// bench / DMC case): no per-step null checks, compile-time tile flush.
// EPW = envs per warp (32, 16 or 8).  The thread-per-env engines are bound by the latency of one env-step's
// dependent instruction chain, so a batch that yields fewer warps than the GPU has scheduler slots (16 384 envs
// = 512 warps for 592 schedulers) runs faster spread over more warps: lanes >= EPW carry no env and only help
// to flush the tiles; fewer envs per warp also means fewer warp-steps pay for a diverged reset.
template <class G, class Ch, class ObsT, int BLOCK, bool ALL, int EPW>
__global__ void __launch_bounds__(BLOCK) k_rollout(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int kRowBytes = G::OBS * (int)sizeof(ObsT);
    constexpr int kWarps = BLOCK / 32;
    constexpr int kTileBytes = kWarps * EPW * kRowBytes;
    if constexpr (G::kSharedBytes > 0) {                      // per-block constant tables of the game
        G::fill_shared(reinterpret_cast<uint8_t *>(smem_raw) + kTileBytes, threadIdx.x, BLOCK);
        __syncthreads();
    }
    const size_t warp_env0 = ((size_t)blockIdx.x * kWarps + wib) * EPW;
    if (warp_env0 >= p.n) return;
    const size_t i = warp_env0 + lane;
    const bool valid = lane < EPW && i < p.n;
    const int nvalid = (int)min((size_t)EPW, p.n - warp_env0);
    ObsT *tile = reinterpret_cast<ObsT *>(smem_raw) + (size_t)wib * EPW * G::OBS;
    ObsT *row = tile + (lane & (EPW - 1)) * G::OBS;
    warp_tile_zero(reinterpret_cast<uint8_t *>(tile), EPW * kRowBytes, lane);
    __syncwarp();

    G g; EnvHeader h; Ch ch; int err = 0;
    g.bind_shared(reinterpret_cast<const uint8_t *>(smem_raw) + kTileBytes);
    // dense mask rows wider than one word (UNO: 61 bytes) are staged per warp too: the rows of a warp are contiguous
    // in global memory, byte stores from the lanes would touch 32 different sectors each
    constexpr bool kStageMask = G::A > 4;
    constexpr int kMaskTile = kStageMask ? ((EPW * G::A + 15) & ~15) : 0;
    uint8_t *mtile = reinterpret_cast<uint8_t *>(smem_raw) + kTileBytes + ((G::kSharedBytes + 15) & ~15) + wib * kMaskTile;
    if constexpr (kStageMask) { warp_tile_zero(mtile, kMaskTile, lane); __syncwarp(); }   // rows OR their edge words into the tile
    constexpr bool kWarpDeal = G::kWarpDeal && Ch::kKind == 0;   // episodes are dealt by the whole warp (UNO, throughput mode)
    constexpr bool kPairEmit = HasPairEmit<G>::value && ALL && EPW == 16 && Ch::kKind == 0 && G::MASK_WORDS == 2;
    bool starts = false;                                         // this lane's env begins an episode at this point
    if (valid) {
        h.load(p.state, p.n, i);
        game_load<G, Ch>(g, p.state + kHeaderWords * p.n, p.n, i);
        ChanceIO<Ch>::open(ch, p, i);
        if (h.episode == 0 || g.over()) {         // never dealt, or a finished episode left by rlc_step without auto reset
            ch.begin_reset(h.k);
            if constexpr (kWarpDeal) { h.episode++; h.t = 0; starts = true; }
            else new_episode(g, ch, h);
        }
    }
    if constexpr (kWarpDeal) g.warp_deal(ch, starts, lane);
    // per-thread output cursors, advanced by one trajectory row (n envs) per step
    uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + warp_env0 * (size_t)kRowBytes;
    const size_t obs_step = p.n * (size_t)kRowBytes;
    const bool full_warp = nvalid == EPW;                     // ALL: aligned rows were checked by the launcher
    size_t rowi = i;
    for (int t = 0; t < p.T; t++, rowi += p.n, o_obs += obs_step) {
        uint32_t m[G::MASK_WORDS];
        starts = false;
        if constexpr (kPairEmit) {                                // lanes 16..31 build half of every row (see encode_obs_pair)
            m[0] = m[1] = 0u;
            if (valid) g.legal(m);
            g.encode_obs_pair(valid, lane, row);
            stage_mask_row_pair<G>(mtile, lane, valid, m);
        } else if (valid) {
            if (ALL || p.t_obs) g.encode_obs(g.player(), h.t == 0, row);
            g.legal(m);
            if constexpr (kStageMask) {
                if (ALL || p.t_mask) {
                    stage_mask_row<G, EPW>(mtile, lane, m);
                }
            }
        }
        __syncwarp();
        if constexpr (kStageMask) {
            uint8_t *gm = reinterpret_cast<uint8_t *>(p.t_mask) + (rowi - lane) * (size_t)G::A;
            if constexpr (ALL) {
                if (full_warp) tile_store_begin<EPW * G::A>(gm, mtile, lane);
                else warp_tile_flush(gm, mtile, nvalid * G::A, lane);
            } else if (p.t_mask) warp_tile_flush(gm, mtile, nvalid * G::A, lane);
        }
        if constexpr (ALL) {
            if (full_warp) tile_store_begin<EPW * kRowBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
            else warp_tile_flush(o_obs, reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
        } else {
            if (p.t_obs) warp_tile_flush(o_obs, reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
        }
        __syncwarp();
        if (valid) {
            if constexpr (!kStageMask) {
                if (ALL || p.t_mask) write_mask_row<G>(reinterpret_cast<uint8_t *>(p.t_mask), rowi, m);
            }
            if (ALL || p.t_player) st_stream(p.t_player + rowi, g.player());
            uint32_t word;
            if constexpr (Ch::kKind == 0) word = ch.begin_step(h.k);
            else word = policy_word_only(p, i, h.k);
            int cnt;
            int a = pick_action<G>(m, word, cnt);
            if constexpr (G::kUsesChain) ch.seed_chain(word, (uint32_t)cnt);
            bool idle = false;
            if constexpr (!ALL) {
                if (p.t_forced) { a = p.t_forced[rowi]; idle = a < 0; }   // replay of a recorded action sequence
            }
            if (ALL || p.t_action) st_stream(p.t_action + rowi, a);
            bool over = false;
            float pay[G::P];
#pragma unroll
            for (int q = 0; q < G::P; q++) pay[q] = 0.f;
            if (!idle) {
                if constexpr (G::kHasApply) {
                    if (ALL || !p.t_forced) g.apply(a, ch, err);      // a was picked from the legal set: no re-validation
                    else g.step(a, ch, err);
                } else g.step(a, ch, err);
                h.t++; h.k++;
                over = g.over();
            }
            if constexpr (!ALL) {
                if (p.tm_row) {                                       // per-seat terminal states into the pool
                    int r = -1;
                    if (over) {
                        r = atomicAdd(p.tm_count, 1);
                        if (r >= p.tm_cap) { r = -1; err |= kErrTerminalPoolFull; }
                    }
                    p.tm_row[rowi] = r;
                    if (r >= 0) {
                        if (p.tm_obs) {                               // this lane's tile row is zero here (flushed above)
                            ObsT *dst = reinterpret_cast<ObsT *>(p.tm_obs) + (size_t)r * (G::P * G::OBS);
                            for (int s = 0; s < G::P; s++) {
                                g.encode_obs(s, false, row);
                                for (int q = 0; q < G::OBS; q++) { dst[s * G::OBS + q] = row[q]; row[q] = (ObsT)0; }
                            }
                        }
                        if (p.tm_mask) {
                            uint32_t mt[G::MASK_WORDS];
                            g.legal(mt);
                            uint8_t *dm = reinterpret_cast<uint8_t *>(p.tm_mask) + (size_t)r * G::A;
                            for (int q = 0; q < G::A; q++) dm[q] = (mt[q >> 5] >> (q & 31)) & 1u;
                        }
                    }
                }
            }
            if (over) {
                g.payoffs(pay);
                if constexpr (kWarpDeal) { h.episode++; h.t = 0; starts = true; }
                else new_episode(g, ch, h);
            }
            if (ALL || p.t_done) p.t_done[rowi] = over ? 1 : 0;
            if (ALL || p.t_payoffs) {
                if constexpr (G::P == 2) st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, make_float2(pay[0], pay[1]));
                else if constexpr (G::P == 4) st_stream(reinterpret_cast<float4 *>(p.t_payoffs) + rowi, make_float4(pay[0], pay[1], pay[2], pay[3]));
                else {
#pragma unroll
                    for (int q = 0; q < G::P; q++) p.t_payoffs[rowi * G::P + q] = pay[q];
                }
            }
        }
        if constexpr (kWarpDeal) g.warp_deal(ch, starts, lane);
        if constexpr (ALL) {                                  // the tiles are free (and zero) again before the next rows
            if (full_warp) {
                if constexpr (kStageMask) tile_store_end<EPW * G::A>(mtile, lane);
                tile_store_end<EPW * kRowBytes>(reinterpret_cast<uint8_t *>(tile), lane);
#if RLC_TMA_FLUSH
                __syncwarp();
#endif
            }
        }
    }
    if (valid) {
        ChanceIO<Ch>::close(ch, p, i);
        h.store(p.state, p.n, i);
        game_store<G, Ch>(g, p.state + kHeaderWords * p.n, p.n, i);
        err |= ch.err;
        if (err && p.err) p.err[i] |= err;
    }
}

// ---- host-side launcher shared by the per-game translation units ----------------------------
enum { kOpReset = 0, kOpStep = 1, kOpObserve = 2, kOpRollout = 3 };

// envs per warp of the fused rollout: 32 unless the batch leaves warp schedulers short of kRolloutWarpsPerScheduler
// warps AND the game gains from fewer envs per warp (G::kRolloutMinEpw; measured on a B200 at 16 384 envs, ms per
// 128-step launch at 32 / 16 / 8: UNO 0.405 / 0.392 / 0.587 (0.658 / 0.585 / 0.730 before the warp-cooperative deal),
// Limit 0.106 / 0.134 / 0.200, no-limit 0.197 / 0.281 / 0.483, Blackjack (65 536 envs) 0.614 / 0.900 / 1.460 -- only
// UNO, whose warps diverge on draws and resets, wins).  RLC_ROLLOUT_EPW forces a value (tests, tuning).  Only the
// Philox throughput mode is specialised, replays always run 32 envs per warp.
constexpr int kRolloutWarpsPerScheduler = 2;
inline int rollout_envs_per_warp(size_t n, int min_epw) {
    static int schedulers = 0;
    if (schedulers == 0) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        schedulers = 4 * sms;
    }
    const char *e = getenv("RLC_ROLLOUT_EPW");
    const int forced = e ? atoi(e) : 0;
    if (forced == 32 || forced == 16 || forced == 8) return forced;
    for (int epw = 32; epw > min_epw; epw >>= 1)
        if (n / (size_t)epw >= (size_t)kRolloutWarpsPerScheduler * schedulers) return epw;
    return min_epw;
}

template <class G, class Ch, class ObsT, int EPW>
cudaError_t launch_rollout(const KParams &p, cudaStream_t stream) {
    constexpr int BLOCK = 64, kWarps = BLOCK / 32;
    constexpr int kRowBytes = G::OBS * (int)sizeof(ObsT);
    const unsigned grid = (unsigned)((p.n + kWarps * EPW - 1) / (kWarps * EPW));
    const size_t smem = (size_t)kWarps * EPW * kRowBytes + ((G::kSharedBytes + 15) & ~15) +
                        (G::A > 4 ? (size_t)kWarps * ((EPW * G::A + 15) & ~15) : 0);
    // fast path: every trajectory stream requested and the obs (and staged mask) rows of a full warp 16-byte aligned
    constexpr bool kTilesAligned = (EPW * kRowBytes) % 16 == 0 && (G::A <= 4 || (EPW * G::A) % 16 == 0);
    const bool all = kTilesAligned && !p.t_forced && !p.tm_row &&
                     p.t_obs && p.t_mask && p.t_action && p.t_player && p.t_done && p.t_payoffs &&
                     ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * (size_t)kRowBytes)) & 15u) == 0 &&
                     (G::A <= 4 || ((reinterpret_cast<uintptr_t>(p.t_mask) | (p.n * (size_t)G::A)) & 15u) == 0);
    cudaError_t e = cudaSuccess;
    if constexpr (kTilesAligned) {
        if (all) {
            auto k = k_rollout<G, Ch, ObsT, BLOCK, true, EPW>;
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) { k<<<grid, BLOCK, smem, stream>>>(p); e = cudaGetLastError(); }
            return e;
        }
    }
    auto k = k_rollout<G, Ch, ObsT, BLOCK, false, EPW>;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) { k<<<grid, BLOCK, smem, stream>>>(p); e = cudaGetLastError(); }
    return e;
}

template <class G, class Ch, class ObsT>
cudaError_t launch_op(int op, const KParams &p, cudaStream_t stream) {
    constexpr int BLOCK = 64;
    if (op == kOpRollout) {
        if constexpr (Ch::kKind == 0) {
            const int epw = rollout_envs_per_warp(p.n, G::kRolloutMinEpw);
            if (epw == 16) return launch_rollout<G, Ch, ObsT, 16>(p, stream);
            if (epw == 8) return launch_rollout<G, Ch, ObsT, 8>(p, stream);
        }
        return launch_rollout<G, Ch, ObsT, 32>(p, stream);
    }
    const unsigned grid = (unsigned)((p.n + BLOCK - 1) / BLOCK);
    const size_t smem = (size_t)BLOCK * G::OBS * sizeof(ObsT) + ((G::kSharedBytes + 15) & ~15);
    cudaError_t e = cudaSuccess;
#define RLC_LAUNCH(KERNEL)                                                                           \
    do {                                                                                             \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) { KERNEL<<<grid, BLOCK, smem, stream>>>(p); e = cudaGetLastError(); }  \
    } while (0)
    switch (op) {
    case kOpReset: RLC_LAUNCH((k_env<G, Ch, ObsT, kModeReset, BLOCK>)); break;
    case kOpStep: RLC_LAUNCH((k_env<G, Ch, ObsT, kModeStep, BLOCK>)); break;
    case kOpObserve: RLC_LAUNCH((k_env<G, Ch, ObsT, kModeObserve, BLOCK>)); break;
    default: e = cudaErrorInvalidValue;
    }
#undef RLC_LAUNCH
    return e;
}

template <class G>
cudaError_t dispatch_game(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t stream) {
    if (obs_dtype == RLC_U8) {
        if (chance == RLC_CHANCE_PHILOX) return launch_op<G, ChancePhilox, uint8_t>(op, p, stream);
        if (chance == RLC_CHANCE_TAPE) return launch_op<G, ChanceTape, uint8_t>(op, p, stream);
        if (chance == RLC_CHANCE_MT19937) return launch_op<G, ChanceMt, uint8_t>(op, p, stream);
    } else if (obs_dtype == RLC_F32) {
        if (chance == RLC_CHANCE_PHILOX) return launch_op<G, ChancePhilox, float>(op, p, stream);
        if (chance == RLC_CHANCE_TAPE) return launch_op<G, ChanceTape, float>(op, p, stream);
        if (chance == RLC_CHANCE_MT19937) return launch_op<G, ChanceMt, float>(op, p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace rlc
