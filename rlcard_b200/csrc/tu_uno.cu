// tu_uno.cu -- kernel instantiations for Uno (one translation unit per game: parallel nvcc)
#include "game_uno.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_uno(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Uno>(op, chance, obs_dtype, p, s);
}
}  // namespace rlc
