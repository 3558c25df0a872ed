// tu_uno.cu -- kernel instantiations for Uno (one translation unit per game: parallel nvcc)
#include "game_uno.cuh"
#include "kernels.cuh"
namespace rlc {
// throughput mode runs the multiset-pile game (UnoBag), the replay modes the ordered-pile game (Uno)
cudaError_t dispatch_uno(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    if (chance == RLC_CHANCE_PHILOX) {
        if (obs_dtype == RLC_U8) return launch_op<UnoBag, ChancePhilox, uint8_t>(op, p, s);
        if (obs_dtype == RLC_F32) return launch_op<UnoBag, ChancePhilox, float>(op, p, s);
        return cudaErrorInvalidValue;
    }
    if (obs_dtype == RLC_U8) {
        if (chance == RLC_CHANCE_TAPE) return launch_op<Uno, ChanceTape, uint8_t>(op, p, s);
        if (chance == RLC_CHANCE_MT19937) return launch_op<Uno, ChanceMt, uint8_t>(op, p, s);
    } else if (obs_dtype == RLC_F32) {
        if (chance == RLC_CHANCE_TAPE) return launch_op<Uno, ChanceTape, float>(op, p, s);
        if (chance == RLC_CHANCE_MT19937) return launch_op<Uno, ChanceMt, float>(op, p, s);
    }
    return cudaErrorInvalidValue;
}

// games/uno/utils.py:86-127 encode_hand + encode_target as a standalone operator: one thread per case
__global__ void k_encode_uno(const uint8_t *hands, const uint8_t *targets, int n, uint8_t *obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    UnoBag g;
#pragma unroll
    for (int p = 0; p < 2; p++) { g.hc[p][0] = g.hc[p][1] = g.hc[p][2] = g.hc[p][3] = 0; g.hw[p] = 0; }
    for (int k = 0; k < 32; k++) { const int code = hands[(size_t)i * 32 + k]; if (code < 60) g.to_hand(0, code); }
    g.tcode = targets[i]; g.tcolor = targets[i] / 15;
    uint8_t *row = obs + (size_t)i * 240;
    for (int k = 0; k < 240; k++) row[k] = 0;
    g.encode_obs(0, false, row);
}
cudaError_t encode_uno(const uint8_t *hands, const uint8_t *targets, int n, uint8_t *obs, cudaStream_t s) {
    k_encode_uno<<<(n + 63) / 64, 64, 0, s>>>(hands, targets, n, obs);
    return cudaGetLastError();
}
}  // namespace rlc
