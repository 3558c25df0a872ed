// tu_leduc.cu -- kernel instantiations for Leduc (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_leduc(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Leduc>(op, chance, obs_dtype, p, s);
}

// leducholdem/judger.py:12-64 + game.py:170-178 as a standalone operator: one thread per case
__global__ void k_judge_leduc(const int32_t *cases, int n, float *payoffs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t *c = cases + (size_t)i * 7;
    Leduc g;
    g.hand0 = c[0]; g.hand1 = c[1]; g.pub = c[2] < 0 ? 0 : c[2]; g.pub_dealt = c[2] >= 0;
    g.chips0 = c[3]; g.chips1 = c[4]; g.fold0 = c[5]; g.fold1 = c[6]; g.rc = 2;
    float out[2];
    g.payoffs(out);
    payoffs[2 * i] = out[0]; payoffs[2 * i + 1] = out[1];
}
cudaError_t judge_leduc(const int32_t *cases, int n, float *payoffs, cudaStream_t s) {
    k_judge_leduc<<<(n + 127) / 128, 128, 0, s>>>(cases, n, payoffs);
    return cudaGetLastError();
}
}  // namespace rlc
