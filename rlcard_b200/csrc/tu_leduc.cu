// tu_leduc.cu -- kernel instantiations for Leduc (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_leduc(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Leduc>(op, chance, obs_dtype, p, s);
}
}  // namespace rlc
