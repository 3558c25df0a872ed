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

cudaError_t compact_trajectory(int game, const rlc_trajectory *t, int obs_dtype, size_t cells, uint32_t *out, cudaStream_t s) {
    const unsigned grid = (unsigned)((cells + 255) / 256);
    const uint8_t *mask = reinterpret_cast<const uint8_t *>(t->mask);
#define RLC_PACK(K, T) K<T><<<grid, 256, 0, s>>>(reinterpret_cast<const T *>(t->obs), mask, t->action, t->player, t->done, t->payoffs, cells, out)
    if (game == RLC_LEDUC) { if (obs_dtype == RLC_F32) RLC_PACK(k_compact_leduc, float); else RLC_PACK(k_compact_leduc, uint8_t); }
    else if (game == RLC_LIMIT) { if (obs_dtype == RLC_F32) RLC_PACK(k_compact_limit, float); else RLC_PACK(k_compact_limit, uint8_t); }
    else return cudaErrorNotSupported;
#undef RLC_PACK
    return cudaGetLastError();
}

}  // namespace rlc
