// tu_compact.cu -- compact host wire format of rollout trajectories (rlc_compact_trajectory).
//
// A dense trajectory row is the reference's exact obs layout: one byte (or float) per obs element, mostly zeros
// (Leduc: 4 ones in 36, Limit: <= 11 ones in 72).  A host consumer behind PCIe pays for every one of those bytes
// (57 B per Leduc env-step), so the end-to-end rate is the bus, not the simulator.  This operator re-encodes a window
// of dense rows into fixed-size records that carry the same information; rlcard_b200/compact.py expands them back into
// the dense arrays on the host (bit-exact, tested), lazily and only for the rows a consumer touches.
//   Leduc  (envs/leducholdem.py:41-71)   1 word : own rank [0:2) | public rank + 1 [2:4) (0 = not dealt) | own chips [4:8) |
//          other chips [8:12) | legal nibble [12:16) | action [16:18) | player [18] | done [19] | payoff0 * 4 as int8 [20:28)
//   Limit  (envs/limitholdem.py:40-71)   3 words: visible-card set bits 0..51 (w0, w1) | w2 = raise counters 4 x 3 bits
//          [0:12) | legal nibble [12:16) | action [16:18) | player [18] | done [19] | payoff0 * 2 as int8 [20:28)
// Both games are zero-sum for two players, so payoff1 = -payoff0.
// The wide games take a warp per cell (the lanes read the dense row coalesced and vote the bits together); the last word
// always carries action [0:16) | player [16:18) | done [19]:
//   UNO (envs/uno.py:24-33)  7 words: hand planes as two bit sets over the 60 (colour, trait) positions -- w0, w1 = bit 0 of the
//          count (positions 0..31, 32..59), w2, w3 = bit 1 -- | w4, w5 legal set (61 bits) | w6 = action [0:6) | target position
//          [6:12) | player [16] | done [19] | winner + 1 [20:22) (payoffs +-1)
//   DouDizhu (envs/doudizhu.py:26-91)  33 words: the 16 (landlord: 14) 54-element thermometer blocks as 15 rank-count nibbles
//          each (2 words per block) | w32 = action [0:15) | player [16:18) | done [19] | landlord won [20] | positions of the two
//          cards-left one-hots [21:26), [26:31).  The 27 472-bit legal mask is NOT sent: it is a function of the row (block 0 =
//          the hand, block 2 = the action to beat) and the action table, and compact.py recomputes it on the host.
//   Scout (envs/scout.py:171-235)  20 words: hand tops / bottoms, table tops / bottoms as 16 value nibbles each (w0..w7) |
//          w8..w14 legal set (204 bits) | w15 hand lengths 4 x 5 bits | w16 own score | w17, w18 payoffs as int16 x 4 |
//          w19 = action [0:8) | player [16:18) | done [19] | forced [20] | consecutive scouts [21:23) | owner [23:26) | table length [26:31)
#include <cstring>
#include "common.cuh"
#include "../../include/rlcard_b200.h"

namespace rlc {

template <class ObsT>
__global__ void k_compact_leduc(const ObsT *obs, const uint8_t *mask, const int32_t *action, const int32_t *player,
                                const uint8_t *done, const float *pay, size_t cells, uint32_t *out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const ObsT *row = obs + c * 36;
    uint32_t own = 0, pub = 0, mine = 0, other = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { own |= row[k] != (ObsT)0 ? (uint32_t)k : 0u; pub |= row[3 + k] != (ObsT)0 ? (uint32_t)(k + 1) : 0u; }
#pragma unroll
    for (int k = 0; k < 15; k++) { mine |= row[6 + k] != (ObsT)0 ? (uint32_t)k : 0u; other |= row[21 + k] != (ObsT)0 ? (uint32_t)k : 0u; }
    const uint32_t m4 = reinterpret_cast<const uint32_t *>(mask)[c];
    const uint32_t nib = (m4 & 1u) | ((m4 >> 7) & 2u) | ((m4 >> 14) & 4u) | ((m4 >> 21) & 8u);
    const int q = (int)(pay[2 * c] * 4.f);
    out[c] = own | (pub << 2) | (mine << 4) | (other << 8) | (nib << 12) | (((uint32_t)action[c] & 3u) << 16) |
             (((uint32_t)player[c] & 1u) << 18) | ((done[c] ? 1u : 0u) << 19) | (((uint32_t)q & 255u) << 20);
}

template <class ObsT>
__global__ void k_compact_limit(const ObsT *obs, const uint8_t *mask, const int32_t *action, const int32_t *player,
                                const uint8_t *done, const float *pay, size_t cells, uint32_t *out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const ObsT *row = obs + c * 72;
    uint64_t cards = 0;
#pragma unroll 4
    for (int k = 0; k < 52; k++) cards |= row[k] != (ObsT)0 ? 1ull << k : 0ull;
    uint32_t rn = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 5; k++) rn |= row[52 + 5 * r + k] != (ObsT)0 ? (uint32_t)k << (3 * r) : 0u;
    const uint32_t m4 = reinterpret_cast<const uint32_t *>(mask)[c];
    const uint32_t nib = (m4 & 1u) | ((m4 >> 7) & 2u) | ((m4 >> 14) & 4u) | ((m4 >> 21) & 8u);
    const int q = (int)(pay[2 * c] * 2.f);
    out[3 * c] = (uint32_t)cards; out[3 * c + 1] = (uint32_t)(cards >> 32);
    out[3 * c + 2] = rn | (nib << 12) | (((uint32_t)action[c] & 3u) << 16) | (((uint32_t)player[c] & 1u) << 18) |
                     ((done[c] ? 1u : 0u) << 19) | (((uint32_t)q & 255u) << 20);
}

template <class ObsT>
__global__ void __launch_bounds__(256) k_compact_uno(const ObsT *obs, const uint8_t *mask, const int32_t *action, const int32_t *player,
                                                     const uint8_t *done, const float *pay, size_t cells, uint32_t *out) {
    const int lane = threadIdx.x & 31;
    const size_t c = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= cells) return;
    const ObsT *row = obs + c * 240;
    const uint8_t *mr = mask + c * 61;
    uint32_t w[7];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int idx = lane + 32 * h;
        int k = 0, tgt = 0;
        if (idx < 60) {
            k = row[60 + idx] != (ObsT)0 ? 1 : (row[120 + idx] != (ObsT)0 ? 2 : 0);
            tgt = row[180 + idx] != (ObsT)0;
        }
        w[h] = __ballot_sync(0xffffffffu, k & 1);
        w[2 + h] = __ballot_sync(0xffffffffu, k >> 1);
        w[4 + h] = __ballot_sync(0xffffffffu, idx < 61 && mr[idx] != 0);
        const uint32_t tb = __ballot_sync(0xffffffffu, tgt);
        if (h == 0) w[6] = tb ? (uint32_t)(__ffs(tb) - 1) : 0u;
        else if (tb) w[6] = 32u + (uint32_t)(__ffs(tb) - 1);
    }
    const float p0 = pay[2 * c];
    w[6] = ((uint32_t)action[c] & 63u) | (w[6] << 6) | (((uint32_t)player[c] & 1u) << 16) | ((done[c] ? 1u : 0u) << 19) |
           ((p0 > 0.f ? 1u : (p0 < 0.f ? 2u : 0u)) << 20);
    if (lane < 7) {
        uint32_t v = w[0];
#pragma unroll
        for (int q = 1; q < 7; q++) v = lane == q ? w[q] : v;
        out[c * 7 + lane] = v;
    }
}

template <class ObsT>
__global__ void __launch_bounds__(256) k_compact_doudizhu(const ObsT *obs, int obs_stride, const int32_t *action, const int32_t *player,
                                                          const uint8_t *done, const float *pay, size_t cells, uint32_t *out) {
    const int lane = threadIdx.x & 31;
    const size_t c = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= cells) return;
    const ObsT *row = obs + c * (size_t)obs_stride;
    const int seat = player[c];
    uint32_t lo = 0, hi = 0, pos = 0;
    if (lane < 16) {                                             // block `lane`: 13 x 4 thermometer + two joker elements
        if (lane < (seat == 0 ? 14 : 16)) {
            const ObsT *b = row + 54 * lane;
            for (int r = 0; r < 13; r++) {
                const uint32_t n = (b[4 * r] != (ObsT)0) + (b[4 * r + 1] != (ObsT)0) + (b[4 * r + 2] != (ObsT)0) + (b[4 * r + 3] != (ObsT)0);
                if (r < 8) lo |= n << (4 * r); else hi |= n << (4 * (r - 8));
            }
            hi |= (b[52] != (ObsT)0 ? 1u : 0u) << 20;
            hi |= (b[53] != (ObsT)0 ? 1u : 0u) << 24;
        }
        out[c * 33 + 2 * lane] = lo; out[c * 33 + 2 * lane + 1] = hi;
    } else if (lane < 18) {                                      // the two cards-left one-hots (landlord 17 + 17, peasant 20 + 17)
        const int first = lane == 16;
        const int base = seat == 0 ? (first ? 756 : 773) : (first ? 864 : 884), size = (seat != 0 && first) ? 20 : 17;
        for (int k = 0; k < size; k++) if (row[base + k] != (ObsT)0) pos = (uint32_t)k;
    }
    const uint32_t pa = __shfl_sync(0xffffffffu, pos, 16), pb = __shfl_sync(0xffffffffu, pos, 17);
    if (lane == 0)
        out[c * 33 + 32] = ((uint32_t)action[c] & 0x7fffu) | (((uint32_t)seat & 3u) << 16) | ((done[c] ? 1u : 0u) << 19) |
                           ((pay[3 * c] > 0.f ? 1u : 0u) << 20) | (pa << 21) | (pb << 26);
}

__global__ void __launch_bounds__(256) k_compact_scout(const float *obs, const uint8_t *mask, const int32_t *action, const int32_t *player,
                                                       const uint8_t *done, const float *pay, size_t cells, uint32_t *out) {
    const int lane = threadIdx.x & 31;
    const size_t c = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= cells) return;
    const float *row = obs + c * 688;
    uint32_t *o = out + c * 20;
    // lane = (plane, slot pair): planes hand tops / hand bottoms / table tops / table bottoms of 16 slots x 10 values
    {
        const int plane = lane >> 3, s0 = (lane & 7) * 2;
        uint32_t nib2 = 0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float *v = row + 160 * plane + 10 * (s0 + q);
            uint32_t val = 0;
            for (int k = 0; k < 10; k++) if (v[k] != 0.f) val = (uint32_t)(k + 1);
            nib2 |= val << (4 * q);
        }
        // eight lanes of a plane hold one byte each of its 64-bit nibble string
        const uint32_t byte_ = nib2 << (8 * (lane & 3));
        uint32_t word = byte_;
        word |= __shfl_xor_sync(0xffffffffu, word, 1); word |= __shfl_xor_sync(0xffffffffu, word, 2);
        if ((lane & 3) == 0) o[2 * plane + ((lane >> 2) & 1)] = word;
    }
    for (int q = 0; q < 7; q++) {                                 // 204 mask bytes -> 7 words
        const int a = 32 * q + lane;
        const uint32_t b = __ballot_sync(0xffffffffu, a < 204 && mask[c * 204 + a] != 0);
        if (lane == 0) o[8 + q] = b;
    }
    if (lane == 0) {
        const float *sc = row + 672;
        uint32_t owner = 4;
        for (int k = 0; k < 5; k++) if (sc[k] != 0.f) owner = (uint32_t)k;
        const uint32_t consec = (uint32_t)__float2int_rn(sc[5] * 3.f), tl = (uint32_t)__float2int_rn(sc[11] * 16.f);
        uint32_t hl = 0;
        for (int k = 0; k < 4; k++) hl |= ((uint32_t)__float2int_rn(sc[6 + k] * 16.f) & 31u) << (5 * k);
        o[15] = hl;
        o[16] = (uint32_t)__float2int_rn(sc[10] * 16.f);
        const float *py = pay + 4 * c;
        o[17] = ((uint32_t)__float2int_rn(py[0]) & 0xffffu) | ((uint32_t)__float2int_rn(py[1]) << 16);
        o[18] = ((uint32_t)__float2int_rn(py[2]) & 0xffffu) | ((uint32_t)__float2int_rn(py[3]) << 16);
        o[19] = ((uint32_t)action[c] & 255u) | (((uint32_t)player[c] & 3u) << 16) | ((done[c] ? 1u : 0u) << 19) |
                ((sc[13] != 0.f ? 1u : 0u) << 20) | ((consec & 3u) << 21) | (owner << 23) | ((tl & 31u) << 26);
    }
}

cudaError_t compact_trajectory(int game, const rlc_trajectory *t, int obs_dtype, size_t cells, uint32_t *out, cudaStream_t s) {
    const unsigned grid = (unsigned)((cells + 255) / 256);
    const uint8_t *mask = reinterpret_cast<const uint8_t *>(t->mask);
#define RLC_PACK(K, T) K<T><<<grid, 256, 0, s>>>(reinterpret_cast<const T *>(t->obs), mask, t->action, t->player, t->done, t->payoffs, cells, out)
    if (game == RLC_LEDUC) { if (obs_dtype == RLC_F32) RLC_PACK(k_compact_leduc, float); else RLC_PACK(k_compact_leduc, uint8_t); }
    else if (game == RLC_LIMIT) { if (obs_dtype == RLC_F32) RLC_PACK(k_compact_limit, float); else RLC_PACK(k_compact_limit, uint8_t); }
    else if (game == RLC_UNO || game == RLC_DOUDIZHU || game == RLC_SCOUT) {
        const unsigned wgrid = (unsigned)((cells + 7) / 8);                 // a warp per cell
        if (game == RLC_UNO) {
            if (obs_dtype == RLC_F32) k_compact_uno<float><<<wgrid, 256, 0, s>>>(reinterpret_cast<const float *>(t->obs), mask, t->action, t->player, t->done, t->payoffs, cells, out);
            else k_compact_uno<uint8_t><<<wgrid, 256, 0, s>>>(reinterpret_cast<const uint8_t *>(t->obs), mask, t->action, t->player, t->done, t->payoffs, cells, out);
        } else if (game == RLC_DOUDIZHU) {
            if (obs_dtype == RLC_F32) k_compact_doudizhu<float><<<wgrid, 256, 0, s>>>(reinterpret_cast<const float *>(t->obs), 912, t->action, t->player, t->done, t->payoffs, cells, out);
            else k_compact_doudizhu<uint8_t><<<wgrid, 256, 0, s>>>(reinterpret_cast<const uint8_t *>(t->obs), 912, t->action, t->player, t->done, t->payoffs, cells, out);
        } else {
            if (obs_dtype != RLC_F32) return cudaErrorNotSupported;
            k_compact_scout<<<wgrid, 256, 0, s>>>(reinterpret_cast<const float *>(t->obs), mask, t->action, t->player, t->done, t->payoffs, cells, out);
        }
    }
    else return cudaErrorNotSupported;
#undef RLC_PACK
    return cudaGetLastError();
}

}  // namespace rlc
