// tu_uno.cu -- kernel instantiations for Uno (one translation unit per game: parallel nvcc)
#include "game_uno.cuh"
#include "kernels.cuh"
namespace rlc {

// ==========================================================================================
// UNO rollout split over two warps per group of EPW envs (throughput mode, every trajectory stream, whole groups):
//   ENV   lane = env: legal set -> policy -> apply (draws, penalties, reshuffles) -> payoffs -> warp-cooperative opening
//         deal; writes the action / done / payoffs streams and publishes an 8-word snapshot per env-step (the acting
//         seat's four colour words + wild lists, target, legal set, player);
//   EMIT  lane = env: rebuilds the 240-element obs row and the 61-byte mask row from the snapshot in the warp's tiles and
//         streams them out (plus the player stream).
// With a thread per env the obs planes (60 one-hot byte stores), the mask staging and the two tile flushes are ~30 % of the
// instructions of an env-step and sit on the same dependent chain as the transition; here they run concurrently with the
// next transition.  Same engine functions (game_uno.cuh), same Philox draws: results equal k_rollout<UnoBag, ...>.
// ==========================================================================================
constexpr int kUnoSnapWords = 8, kUnoSnapDepth = 4;

template <class ObsT, int EPW>
__global__ void __launch_bounds__(64) k_rollout_uno_ee(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    using G = UnoBag;
    constexpr int kRowBytes = G::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = EPW * kRowBytes, kMaskTile = (EPW * G::A + 15) & ~15;
    uint8_t *sm = reinterpret_cast<uint8_t *>(smem_raw);
    ObsT *tile = reinterpret_cast<ObsT *>(sm);
    uint8_t *mtile = sm + kTileBytes;
    volatile uint32_t *snap = reinterpret_cast<volatile uint32_t *>(sm + kTileBytes + kMaskTile);          // [depth][words][32]
    volatile uint32_t *pub = snap + kUnoSnapDepth * kUnoSnapWords * 32, *emitted = pub + 1;
    const int lane = threadIdx.x & 31;
    const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) & 1);      // 0 ENV, 1 EMIT
    const size_t env0 = (size_t)blockIdx.x * EPW;
    const size_t i = env0 + lane;
    const bool valid = lane < EPW;                                       // the launcher guarantees n % EPW == 0
    if (threadIdx.x == 0) { *pub = 0; *emitted = 0; }
    __syncthreads();

    if (role == 1) {                                                     // ---- EMIT
        ObsT *row = tile + (lane & (EPW - 1)) * G::OBS;
        warp_tile_zero(reinterpret_cast<uint8_t *>(tile), kTileBytes, lane);
        warp_tile_zero(mtile, kMaskTile, lane);
        __syncwarp();
        uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + env0 * (size_t)kRowBytes;
        uint8_t *o_mask = reinterpret_cast<uint8_t *>(p.t_mask) + env0 * (size_t)G::A;
        const size_t obs_step = p.n * (size_t)kRowBytes, mask_step = p.n * (size_t)G::A;
        size_t rowi = i;
        for (int t = 0; t < p.T; t++, rowi += p.n, o_obs += obs_step, o_mask += mask_step) {
            while (*pub <= (uint32_t)t) { }
            __threadfence_block();
            if (valid) {
                const volatile uint32_t *sp = snap + (t % kUnoSnapDepth) * (kUnoSnapWords * 32) + lane;
                G e;
                e.hc[0][0] = sp[0]; e.hc[0][1] = sp[32]; e.hc[0][2] = sp[64]; e.hc[0][3] = sp[96]; e.hw[0] = sp[128];
                const uint32_t meta = sp[160];
                const uint32_t m[2] = { sp[192], sp[224] };
                e.tcode = (int)(meta & 63u);
                e.encode_obs(0, false, row);                             // seat's words were published as seat 0
                stage_mask_row<G, EPW>(mtile, lane, m);
                st_stream(p.t_player + rowi, (int)((meta >> 8) & 1u));
            }
            __syncwarp();
            if (lane == 0) *emitted = (uint32_t)t + 1u;                  // every lane has consumed its snapshot
            tile_store_begin<EPW * G::A>(o_mask, mtile, lane);
            tile_store_begin<kTileBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
            tile_store_end<EPW * G::A>(mtile, lane);
            tile_store_end<kTileBytes>(reinterpret_cast<uint8_t *>(tile), lane);
            __syncwarp();
        }
        return;
    }

    // ---- ENV
    G g; EnvHeader h; ChancePhilox ch; int err = 0;
    bool starts = false;
    if (valid) {
        h.load(p.state, p.n, i);
        g.load(p.state + kHeaderWords * p.n, p.n, i);
        ch.init(p.seed, p.env_id_base + (uint32_t)i);
        if (h.episode == 0 || g.over()) { ch.begin_reset(h.k); h.episode++; h.t = 0; starts = true; }
    }
    g.warp_deal(ch, starts, lane);
    size_t rowi = i;
    for (int t = 0; t < p.T; t++, rowi += p.n) {
        uint32_t m[2] = {0u, 0u};
        starts = false;
        if (valid) g.legal(m);
        {   // publish the snapshot of this env-step
            while ((uint32_t)t - *emitted >= (uint32_t)kUnoSnapDepth) { }
            if (valid) {
                volatile uint32_t *sp = snap + (t % kUnoSnapDepth) * (kUnoSnapWords * 32) + lane;
                const int seat = g.cur;
                sp[0] = g.hcp(seat, 0); sp[32] = g.hcp(seat, 1); sp[64] = g.hcp(seat, 2); sp[96] = g.hcp(seat, 3); sp[128] = g.hwp(seat);
                sp[160] = (uint32_t)g.tcode | ((uint32_t)seat << 8);
                sp[192] = m[0]; sp[224] = m[1];
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *pub = (uint32_t)t + 1u;
        }
        if (valid) {
            const uint32_t word = ch.begin_step(h.k);
            int cnt;
            const int a = pick_action<G>(m, word, cnt);
            st_stream(p.t_action + rowi, a);
            g.apply(a, ch, err);
            h.t++; h.k++;
            const bool over = g.over();
            float pay[2] = {0.f, 0.f};
            if (over) { g.payoffs(pay); h.episode++; h.t = 0; starts = true; }
            p.t_done[rowi] = over ? 1 : 0;
            st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, make_float2(pay[0], pay[1]));
        }
        g.warp_deal(ch, starts, lane);
    }
    if (valid) {
        h.store(p.state, p.n, i);
        g.store(p.state + kHeaderWords * p.n, p.n, i);
        err |= ch.err;
        if (err && p.err) p.err[i] |= err;
    }
}

template <class ObsT, int EPW>
static cudaError_t launch_uno_ee(const KParams &p, cudaStream_t s) {
    constexpr int kRowBytes = UnoBag::OBS * (int)sizeof(ObsT);
    const size_t smem = (size_t)EPW * kRowBytes + ((EPW * UnoBag::A + 15) & ~15) + kUnoSnapDepth * kUnoSnapWords * 32 * 4 + 16;
    k_rollout_uno_ee<ObsT, EPW><<<(unsigned)(p.n / EPW), 64, smem, s>>>(p);
    return cudaGetLastError();
}

// ==========================================================================================
// k_rollout_uno_pipe: the same ENV / EMIT split, handed over in CHUNKS.  The ENV warp appends its 8-word snapshots to a
// double-buffered chunk of kUnoChunk env-steps and the EMIT warp consumes whole chunks; the two meet at named barriers
// (bar.arrive / bar.sync, ids 1..4 = full[2], free[2]) once per chunk instead of polling a ring, fencing and publishing a
// counter in every env-step (what made k_rollout_uno_ee slower than the generic kernel, profiles/r02_uno_ee.md).
// Same engine functions, same Philox draws, same trajectory.
// ==========================================================================================
constexpr int kUnoChunk = 8;
__device__ __forceinline__ void uno_named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void uno_named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

template <class ObsT, int EPW>
__global__ void __launch_bounds__(64) k_rollout_uno_pipe(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    using G = UnoBag;
    constexpr int K = kUnoChunk;
    constexpr int kRowBytes = G::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = EPW * kRowBytes, kMaskTile = (EPW * G::A + 15) & ~15;
    uint8_t *sm = reinterpret_cast<uint8_t *>(smem_raw);
    ObsT *tile = reinterpret_cast<ObsT *>(sm);
    uint8_t *mtile = sm + kTileBytes;
    uint32_t *snap = reinterpret_cast<uint32_t *>(sm + kTileBytes + kMaskTile);          // [2][K][words][32]
    const int lane = threadIdx.x & 31;
    const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) & 1);      // 0 ENV, 1 EMIT
    const size_t env0 = (size_t)blockIdx.x * EPW;
    const size_t i = env0 + lane;
    const bool valid = lane < EPW;                                       // the launcher guarantees n % EPW == 0
    const int nchunks = (p.T + K - 1) / K;

    if (role == 1) {                                                     // ---- EMIT
        ObsT *row = tile + (lane & (EPW - 1)) * G::OBS;
        warp_tile_zero(reinterpret_cast<uint8_t *>(tile), kTileBytes, lane);
        warp_tile_zero(mtile, kMaskTile, lane);
        __syncwarp();
        uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + env0 * (size_t)kRowBytes;
        uint8_t *o_mask = reinterpret_cast<uint8_t *>(p.t_mask) + env0 * (size_t)G::A;
        const size_t obs_step = p.n * (size_t)kRowBytes, mask_step = p.n * (size_t)G::A;
        size_t rowi = i;
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1, steps = min(K, p.T - c * K);
            uno_named_sync(1 + b, 64);                                   // the chunk's snapshots are written
            for (int s = 0; s < steps; s++, rowi += p.n, o_obs += obs_step, o_mask += mask_step) {
                if (valid) {
                    const uint32_t *sp = snap + ((b * K + s) * kUnoSnapWords) * 32 + lane;
                    G e;
                    e.hc[0][0] = sp[0]; e.hc[0][1] = sp[32]; e.hc[0][2] = sp[64]; e.hc[0][3] = sp[96]; e.hw[0] = sp[128];
                    const uint32_t meta = sp[160];
                    const uint32_t m[2] = { sp[192], sp[224] };
                    e.tcode = (int)(meta & 63u);
                    e.encode_obs(0, false, row);                         // seat's words were published as seat 0
                    stage_mask_row<G, EPW>(mtile, lane, m);
                    st_stream(p.t_player + rowi, (int)((meta >> 8) & 1u));
                }
                __syncwarp();
                tile_store_begin<EPW * G::A>(o_mask, mtile, lane);
                tile_store_begin<kTileBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
                tile_store_end<EPW * G::A>(mtile, lane);
                tile_store_end<kTileBytes>(reinterpret_cast<uint8_t *>(tile), lane);
                __syncwarp();
            }
            if (c + 2 < nchunks) { __threadfence_block(); uno_named_arrive(3 + b, 64); }   // the buffer may be rewritten
        }
        return;
    }

    // ---- ENV
    G g; EnvHeader h; ChancePhilox ch; int err = 0;
    bool starts = false;
    if (valid) {
        h.load(p.state, p.n, i);
        g.load(p.state + kHeaderWords * p.n, p.n, i);
        ch.init(p.seed, p.env_id_base + (uint32_t)i);
        if (h.episode == 0 || g.over()) { ch.begin_reset(h.k); h.episode++; h.t = 0; starts = true; }
    }
    g.warp_deal(ch, starts, lane);
    size_t rowi = i;
    for (int c = 0; c < nchunks; c++) {
        const int b = c & 1, steps = min(K, p.T - c * K);
        if (c >= 2) uno_named_sync(3 + b, 64);                           // EMIT is done with this buffer
        for (int s = 0; s < steps; s++, rowi += p.n) {
            uint32_t m[2] = {0u, 0u};
            starts = false;
            if (valid) {
                g.legal(m);
                uint32_t *sp = snap + ((b * K + s) * kUnoSnapWords) * 32 + lane;
                const int seat = g.cur;
                sp[0] = g.hcp(seat, 0); sp[32] = g.hcp(seat, 1); sp[64] = g.hcp(seat, 2); sp[96] = g.hcp(seat, 3); sp[128] = g.hwp(seat);
                sp[160] = (uint32_t)g.tcode | ((uint32_t)seat << 8);
                sp[192] = m[0]; sp[224] = m[1];
                const uint32_t word = ch.begin_step(h.k);
                int cnt;
                const int a = pick_action<G>(m, word, cnt);
                st_stream(p.t_action + rowi, a);
                g.apply(a, ch, err);
                h.t++; h.k++;
                const bool over = g.over();
                float pay[2] = {0.f, 0.f};
                if (over) { g.payoffs(pay); h.episode++; h.t = 0; starts = true; }
                p.t_done[rowi] = over ? 1 : 0;
                st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, make_float2(pay[0], pay[1]));
            }
            g.warp_deal(ch, starts, lane);
        }
        __threadfence_block();
        uno_named_arrive(1 + b, 64);
    }
    if (valid) {
        h.store(p.state, p.n, i);
        g.store(p.state + kHeaderWords * p.n, p.n, i);
        err |= ch.err;
        if (err && p.err) p.err[i] |= err;
    }
}

template <class ObsT, int EPW>
static cudaError_t launch_uno_pipe(const KParams &p, cudaStream_t s) {
    constexpr int kRowBytes = UnoBag::OBS * (int)sizeof(ObsT);
    const size_t smem = (size_t)EPW * kRowBytes + ((EPW * UnoBag::A + 15) & ~15) + 2 * kUnoChunk * kUnoSnapWords * 32 * 4;
    k_rollout_uno_pipe<ObsT, EPW><<<(unsigned)(p.n / EPW), 64, smem, s>>>(p);
    return cudaGetLastError();
}

// throughput mode runs the multiset-pile game (UnoBag), the replay modes the ordered-pile game (Uno)
cudaError_t dispatch_uno(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    if (chance == RLC_CHANCE_PHILOX) {
        const char *pp = getenv("RLC_UNO_PIPE");                         // 16 / 32 = chunked env/emit kernel with that many envs per group
        const int ppw = pp ? atoi(pp) : 0;
        if ((ppw == 16 || ppw == 32) && op == kOpRollout && !(p.flags & kFlagNoFsm) && p.n % ppw == 0 && p.T > 0 && p.t_obs && p.t_mask &&
            p.t_action && p.t_player && p.t_done && p.t_payoffs && obs_dtype == RLC_U8 &&
            ((reinterpret_cast<uintptr_t>(p.t_obs) | reinterpret_cast<uintptr_t>(p.t_mask)) & 15u) == 0) {
            if (ppw == 16) return launch_uno_pipe<uint8_t, 16>(p, s);
            return launch_uno_pipe<uint8_t, 32>(p, s);
        }
        const char *ee = getenv("RLC_UNO_EE");                           // 0 = generic kernel, 16 / 32 = env/emit kernel with that many envs per group
        const int epw = ee ? atoi(ee) : 0;
        if ((epw == 16 || epw == 32) && op == kOpRollout && !(p.flags & kFlagNoFsm) && p.n % epw == 0 && p.t_obs && p.t_mask &&
            p.t_action && p.t_player && p.t_done && p.t_payoffs && obs_dtype == RLC_U8 &&
            ((reinterpret_cast<uintptr_t>(p.t_obs) | reinterpret_cast<uintptr_t>(p.t_mask)) & 15u) == 0) {
            if (epw == 16) return launch_uno_ee<uint8_t, 16>(p, s);
            return launch_uno_ee<uint8_t, 32>(p, s);
        }
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
