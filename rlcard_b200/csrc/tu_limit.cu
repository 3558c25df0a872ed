// tu_limit.cu -- kernel instantiations for Limit (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_limit(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Limit>(op, chance, obs_dtype, p, s);
}
}  // namespace rlc
