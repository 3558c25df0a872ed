// tu_nolimit.cu -- kernel instantiations for no-limit hold'em (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_nolimit(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<NoLimit>(op, chance, obs_dtype, p, s);
}
}  // namespace rlc
