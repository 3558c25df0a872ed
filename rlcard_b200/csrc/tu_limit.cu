// tu_limit.cu -- kernel instantiations for Limit (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_limit(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Limit>(op, chance, obs_dtype, p, s);
}

// limitholdem/utils.py:526-569 compare_hands as a standalone operator: one thread per case
__global__ void k_judge_holdem(const uint8_t *cards, int n, int P, uint8_t *winners) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t str[RLC_MAX_PLAYERS], best = 0;
    for (int p = 0; p < P; p++) {
        const uint8_t *c = cards + ((size_t)i * P + p) * 7;
        str[p] = 0;
        if (c[0] != 255) {
            const int h[7] = { c[0], c[1], c[2], c[3], c[4], c[5], c[6] };
            str[p] = 1u + holdem_strength7(h);
        }
        best = max(best, str[p]);
    }
    for (int p = 0; p < P; p++) winners[(size_t)i * P + p] = (str[p] != 0 && str[p] == best) ? 1 : 0;
}
cudaError_t judge_holdem(const uint8_t *cards, int n, int P, uint8_t *winners, cudaStream_t s) {
    if (P < 1 || P > RLC_MAX_PLAYERS) return cudaErrorInvalidValue;
    k_judge_holdem<<<(n + 127) / 128, 128, 0, s>>>(cards, n, P, winners);
    return cudaGetLastError();
}
}  // namespace rlc
