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
}  // namespace rlc
