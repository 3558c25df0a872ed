// tu_scout.cu -- kernel instantiations for Scout (warp per env; obs are float32 only)
#include "game_scout.cuh"
namespace rlc {
cudaError_t dispatch_scout(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    if (obs_dtype != RLC_F32) return cudaErrorInvalidValue;
    return dispatch_wgame<Scout, float>(op, chance, p, s);
}
}  // namespace rlc
