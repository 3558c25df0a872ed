// tu_blackjack.cu -- kernel instantiations for Blackjack (one translation unit per game: parallel nvcc)
#include "game_blackjack.cuh"
#include "kernels.cuh"
namespace rlc {
cudaError_t dispatch_blackjack(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    return dispatch_game<Blackjack>(op, chance, obs_dtype, p, s);
}
}  // namespace rlc
