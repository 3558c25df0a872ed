// tu_leduc.cu -- kernel instantiations for Leduc (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
#include <mutex>
#include <type_traits>
namespace rlc {

// ==========================================================================================
// Table-driven Leduc rollout.  Everything about a Leduc state except the three cards is one of < 256
// "betting states" (round, pointer, raise counters, chips, folds); the transition function over them is
// tabulated ONCE per device by running the register-level engine (Leduc::step / legal / payoffs / encode_obs,
// game_poker.cuh) over every reachable state, so table == engine by construction.  The rollout then costs two
// 128-bit shared-memory lookups per env-step instead of the branchy engine.  One entry (uint4) per state:
//   .x  the packed state word (game_poker.cuh layout) with the card bits [0:7) replaced by legal mask [0:4) | is_over [4]
//   .y  running state: id of the next state after the k-th legal action (k = 0..3 in ascending id order, a byte each)
//       terminal state: seat 0's payoff in quarter chips (int8) when the cards say seat 0 wins / seat 1 wins / tie
//       (seat 1 gets the negative: Leduc is zero-sum, checked while tabulating; a fold decides all three alike)
//   .z  the id of the k-th legal action (a byte each)
//   .w  where the obs row of the acting seat has its ones (envs/leducholdem.py:41-71): byte 0 = 6 + own chips,
//       byte 1 = 21 + opponent's chips, byte 2 = bit offset of the own card in `cards`, byte 3 = public card dealt
// State ids 0 / 1 are the first states of an episode with seat 0 / 1 as small blind.
// The cards live in one register: hand0 | hand1 << 2 | public << 4 | showdown outcome << 6 (0 seat 0 wins, 1 seat 1
// wins, 2 tie -- judger.py:12-64 evaluated by Leduc::payoffs when the deal table is filled).
// ==========================================================================================
constexpr int kFsmMax = 256;
static uint4 *g_fsm[64];
static std::mutex g_fsm_mu;

__device__ __forceinline__ uint32_t fsm_key(const Leduc &g) {
    return (g.chips0 << 7) | (g.chips1 << 11) | (g.r.raised0 << 15) | (g.r.raised1 << 19) | (g.r.have_raised << 23) |
           (g.r.not_raise_num << 25) | (g.r.pointer << 27) | (g.rc << 28) | (g.fold0 << 30) | ((uint32_t)g.fold1 << 31);
}
__device__ __forceinline__ void fsm_unkey(Leduc &g, uint32_t w) {
    g.hand0 = g.hand1 = g.pub = 0; g.pub_dealt = 0;
    g.chips0 = bf_get(w, 7, 4); g.chips1 = bf_get(w, 11, 4); g.r.raised0 = bf_get(w, 15, 4); g.r.raised1 = bf_get(w, 19, 4);
    g.r.have_raised = bf_get(w, 23, 2); g.r.not_raise_num = bf_get(w, 25, 2); g.r.pointer = bf_get(w, 27, 1);
    g.rc = bf_get(w, 28, 2); g.fold0 = bf_get(w, 30, 1); g.fold1 = bf_get(w, 31, 1);
}
// showdown outcome of a deal (hand0 | hand1 << 2 | public << 4), by the engine's judger on an even pot
__device__ __forceinline__ uint32_t leduc_showdown_code(uint32_t cards) {
    Leduc g; fsm_unkey(g, 0);
    g.hand0 = cards & 3u; g.hand1 = (cards >> 2) & 3u; g.pub = (cards >> 4) & 3u; g.pub_dealt = 1; g.rc = 2;
    g.chips0 = g.chips1 = 1;
    float out[2];
    g.payoffs(out);
    return out[0] > 0.f ? 0u : (out[0] < 0.f ? 1u : 2u);
}
// single thread: breadth-first closure of the two initial states under Leduc::step
__global__ void k_leduc_build_fsm(uint4 *tab, int *count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    __shared__ uint32_t keys[kFsmMax];
    int n = 0;
    for (int sb = 0; sb < 2; sb++) {                         // game.py:74-95 with the blind draw fixed
        Leduc g; fsm_unkey(g, 0);
        g.chips0 = sb == 0 ? 1 : 2; g.chips1 = sb == 0 ? 2 : 1;
        g.r.start(sb, g.chips0, g.chips1);
        keys[n++] = fsm_key(g);
    }
    ChancePhilox none; none.init(0, 0);
    for (int i = 0; i < n; i++) {
        Leduc g; fsm_unkey(g, keys[i]);
        uint32_t m[1]; g.legal(m);
        uint32_t x = keys[i] | (m[0] & 15u) | (g.over() ? 16u : 0u), y = 0, z = 0, w = 0;
        if (!g.over()) {
            int kth = 0;
            for (int a = 0; a < 4; a++) {
                if (!((m[0] >> a) & 1u)) continue;
                Leduc h; fsm_unkey(h, keys[i]);
                int err = 0;
                h.step(a, none, err);
                const uint32_t k2 = fsm_key(h);
                int j = 0;
                while (j < n && keys[j] != k2) j++;
                if (j == n) { if (n >= kFsmMax) { *count = -1; return; } keys[n++] = k2; }
                y |= (uint32_t)j << (8 * kth);
                z |= (uint32_t)a << (8 * kth);
                kth++;
            }
            // where encode_obs puts the ones of the acting seat's row: probe it with distinct cards (own 1, other 2, public 0)
            const int seat = g.player();
            g.hand0 = seat ? 2 : 1; g.hand1 = seat ? 1 : 2; g.pub = 0; g.pub_dealt = g.rc != 0;
            uint8_t row[Leduc::OBS];
            for (int k = 0; k < Leduc::OBS; k++) row[k] = 0;
            g.encode_obs(seat, false, row);
            int mine = -1, other = -1;
            for (int k = 6; k < 21; k++) if (row[k]) mine = k;
            for (int k = 21; k < 36; k++) if (row[k]) other = k;
            if (mine < 0 || other < 0 || !row[1] || (row[3] != 0) != (g.rc != 0)) { *count = -2; return; }
            w = (uint32_t)mine | ((uint32_t)other << 8) | ((uint32_t)(2 * seat) << 16) | ((g.rc != 0 ? 1u : 0u) << 24);
        } else {
            for (int code = 0; code < 3; code++) {           // cards that make the judger say: seat 0 / seat 1 / tie
                g.hand0 = code == 0 ? 1 : 0; g.hand1 = code == 1 ? 1 : 0; g.pub = 1; g.pub_dealt = 1;
                float out[2];
                g.payoffs(out);
                const int q0 = (int)(out[0] * 4.f), q1 = (int)(out[1] * 4.f);
                if (q0 != -q1 || (float)q0 * 0.25f != out[0]) { *count = -3; return; }
                y |= ((uint32_t)q0 & 255u) << (8 * code);
            }
        }
        tab[i] = make_uint4(x, y, z, w);
    }
    *count = n;
}
// Built once per device by rlc_upload_tables(RLC_LEDUC, device, NULL, 0) (VecEnv does it at construction): an explicit,
// synchronous initialisation, so that rlc_rollout_random itself never allocates, never touches another stream and stays
// capturable in a CUDA graph.  Until then the rollout runs the generic register engine (same results).
cudaError_t leduc_init(int device) {
    if (device < 0 || device >= 64) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(g_fsm_mu);
    if (g_fsm[device]) return cudaSuccess;
    int prev = 0; cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    uint4 *tab = nullptr; int *cnt = nullptr, h = 0;
    e = cudaMalloc(&tab, sizeof(uint4) * kFsmMax);
    if (e == cudaSuccess) e = cudaMalloc(&cnt, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(tab, 0, sizeof(uint4) * kFsmMax);
    if (e == cudaSuccess) { k_leduc_build_fsm<<<1, 32>>>(tab, cnt); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(&h, cnt, sizeof h, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h <= 0) e = cudaErrorUnknown;
    if (cnt) cudaFree(cnt);
    if (e == cudaSuccess) g_fsm[device] = tab;
    else if (tab) cudaFree(tab);
    cudaSetDevice(prev);
    return e;
}
static const uint4 *leduc_fsm_on_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_fsm_mu);
    return g_fsm[dev];
}

// byte k (k < 4) of w, zero extended / sign extended: one PRMT each
__device__ __forceinline__ uint32_t byte_of(uint32_t w, uint32_t k) { return __byte_perm(w, 0u, 0x4440u | k); }
__device__ __forceinline__ int sbyte_of(uint32_t w, uint32_t k) {            // prmt's sign-replicate selectors (the
    uint32_t r;                                                              // __byte_perm intrinsic masks them off)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(0x8880u | (k * 0x1111u)));
    return (int)r;
}

// The Env.run loop with random agents (env.py:120-169) over the tabulated engine: same trajectory, state
// words and Philox draws as k_rollout<Leduc, ChancePhilox, ObsT, 64, true, 32>.
template <class ObsT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_rollout_leduc_fsm(const KParams p, const uint4 *__restrict__ gtab) {
    extern __shared__ uint4 smem_raw[];
    constexpr int kRowBytes = Leduc::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = BLOCK * kRowBytes;
    uint4 *stab = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(smem_raw) + kTileBytes);
    uint8_t *lut = reinterpret_cast<uint8_t *>(stab + kFsmMax);
    for (int j = threadIdx.x; j < kFsmMax; j += BLOCK) stab[j] = gtab[j];
    for (int x = threadIdx.x; x < 120; x += BLOCK) {         // deal table with the showdown outcome on top
        const uint32_t c = Leduc::deal_code((uint32_t)x);
        lut[x] = (uint8_t)(c | (leduc_showdown_code(c) << 6));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp_env0 = ((size_t)blockIdx.x * (BLOCK / 32) + wib) * 32;
    if (warp_env0 >= p.n) return;
    const size_t i = warp_env0 + lane;
    const bool valid = i < p.n;
    const int nvalid = (int)min((size_t)32, p.n - warp_env0);
    ObsT *tile = reinterpret_cast<ObsT *>(smem_raw) + (size_t)wib * 32 * Leduc::OBS;
    ObsT *row = tile + lane * Leduc::OBS;
    warp_tile_zero(reinterpret_cast<uint8_t *>(tile), 32 * kRowBytes, lane);
    __syncwarp();

    EnvHeader h; ChancePhilox ch;
    uint32_t cards = 0, sid = 0;           // hand0 | hand1 << 2 | public << 4 | showdown outcome << 6 ; betting-state id
    if (valid) {
        h.load(p.state, p.n, i);
        ch.init(p.seed, p.env_id_base + (uint32_t)i);
        if (h.episode == 0) {              // first deal: a reset outside a step
            ch.begin_reset(h.k);
            cards = lut[ch.chain(120u)];
            sid = ch.chain(2u);
            h.episode = 1; h.t = 0;
        } else {
            const uint32_t w = p.state[kHeaderWords * p.n + i];
            cards = w & 63u;
            cards |= leduc_showdown_code(cards) << 6;
            const uint32_t key = w & ~127u;
            for (int j = 0; j < kFsmMax; j++) if ((stab[j].x & ~127u) == key) { sid = (uint32_t)j; break; }
            if ((stab[sid].x >> 4) & 1u) {   // a finished episode was left in the state (rlc_step without auto reset): deal
                ch.begin_reset(h.k);         // the next one, exactly like the generic kernel
                cards = lut[ch.chain(120u)];
                sid = ch.chain(2u);
                h.episode++; h.t = 0;
            }
        }
    }
    uint4 e = stab[sid];
    uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + warp_env0 * (size_t)kRowBytes;
    const size_t obs_step = p.n * (size_t)kRowBytes;
    const bool full_warp = nvalid == 32;
    size_t rowi = i;
    // the step loop is instantiated twice (full warp: compile-time tile flush without a branch, so that the flush and the
    // transition form one straight-line region the scheduler can interleave; ragged last warp: generic flush)
    auto run = [&](auto full_c) {
    constexpr bool kFull = decltype(full_c)::value;
    const bool live = kFull || valid;                      // a full warp has no idle lane: no branch around the step
    for (int t = 0; t < p.T; t++, rowi += p.n, o_obs += obs_step) {
        const uint32_t legal = e.x & 15u;
        if (live) {                        // envs/leducholdem.py:41-71, positions tabulated from Leduc::encode_obs
            row[(cards >> byte_of(e.w, 2)) & 3u] = (ObsT)1;
            if (e.w >> 24) row[3 + ((cards >> 4) & 3u)] = (ObsT)1;
            row[byte_of(e.w, 0)] = (ObsT)1;
            row[byte_of(e.w, 1)] = (ObsT)1;
        }
        __syncwarp();
        if constexpr (kFull) tile_store_begin<32 * kRowBytes, true>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
        else warp_tile_flush(o_obs, reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
        // no barrier here: the transition below does not touch the tile, so its table look-ups overlap the flush; the
        // barrier that orders the flush's zeroing before the next row is written sits at the end of the step
        if (live) {
            st_stream(reinterpret_cast<uint32_t *>(p.t_mask) + rowi, (legal * 0x00204081u) & 0x01010101u);   // bit a -> byte a
            st_stream(p.t_player + rowi, (int)((e.x >> 27) & 1u));
            const uint32_t word = ch.begin_step(h.k);
            const uint32_t cnt = (uint32_t)__popc(legal);
            const uint32_t kth = __umulhi(word, cnt);        // uniform over the legal ids, ascending
            ch.seed_chain(word, cnt);
            st_stream(p.t_action + rowi, (int)byte_of(e.z, kth));
            sid = byte_of(e.y, kth);
            e = stab[sid];
            h.t++; h.k++;
            const bool over = (e.x >> 4) & 1u;
            float2 pay = make_float2(0.f, 0.f);
            if (over) {                    // judger.py:12-64 through the tabulated quarter-chip payoffs
                const float q0 = (float)sbyte_of(e.y, cards >> 6) * 0.25f;
                pay = make_float2(q0, -q0);
                cards = lut[ch.chain(120u)];                 // game.py:46-95: the next episode, same step's draws
                sid = ch.chain(2u);
                e = stab[sid];
                h.episode++; h.t = 0;
            }
            p.t_done[rowi] = over ? 1 : 0;
            st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, pay);
        }
        if constexpr (kFull) tile_store_end<32 * kRowBytes, true>(reinterpret_cast<uint8_t *>(tile), lane);
        __syncwarp();
    }
    };
    if (full_warp) run(std::true_type{}); else run(std::false_type{});
    if (valid) {
        h.store(p.state, p.n, i);
        const uint32_t rc = (e.x >> 28) & 3u;
        p.state[kHeaderWords * p.n + i] = (e.x & ~127u) | (cards & 63u) | (rc ? 64u : 0u);
    }
}

// Block size of the tabulated rollout.  The kernel streams 57 bytes per env-step and is bound by the memory pipeline, and it runs
// fastest when every SM hosts the SAME number of warps, a multiple of four (one per scheduler), in ONE block: the warps of a
// block start together and stay within a few steps of each other, so at any moment the whole GPU writes the same few trajectory
// rows.  Measured on a B200, 65 536 envs x 128 steps (tools/leduc_blk.sh): blocks of 64 / 128 / 256 / 512 / 1024 threads:
// 0.0901 / 0.0923 / 0.0872 / 0.0830 / 0.156 ms on one box, 0.0916 (64) vs 0.0752 (512) on another (profiles/r02_leduc_blocks.md; a
// block barrier every 4..32 steps to realign the warps gains nothing at 512 threads and 6 % at 256).  So: warps per block = the
// per-SM share of the batch rounded up to a multiple of four, at most 32; batches beyond 148 x 32 warps run in waves of
// 1024-thread blocks.
template <class ObsT, int BLOCK>
static cudaError_t launch_leduc_fsm_block(const KParams &p, const uint4 *tab, cudaStream_t s) {
    const size_t smem = (size_t)BLOCK * Leduc::OBS * sizeof(ObsT) + sizeof(uint4) * kFsmMax + 128;
    auto k = k_rollout_leduc_fsm<ObsT, BLOCK>;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)((p.n + BLOCK - 1) / BLOCK), BLOCK, smem, s>>>(p, tab);
    return cudaGetLastError();
}
template <class ObsT>
static cudaError_t launch_leduc_fsm(const KParams &p, const uint4 *tab, cudaStream_t s) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms = v > 0 ? v : 148;
    }
    const size_t warps = (p.n + 31) / 32;
    size_t w = (warps + (size_t)sms - 1) / (size_t)sms;                   // per-SM share
    w = (w + 3) / 4 * 4;
    int block = (int)(w > 32 ? 32 : w) * 32;
    if (warps <= 2 * (size_t)sms) block = 64;                             // tiny batches: spread them
    const char *bs = getenv("RLC_LEDUC_BLOCK");                           // tuning override (a multiple of 128, or 64)
    if (bs && atoi(bs) > 0) block = atoi(bs);
    switch (block) {
    case 128: return launch_leduc_fsm_block<ObsT, 128>(p, tab, s);
    case 256: return launch_leduc_fsm_block<ObsT, 256>(p, tab, s);
    case 384: return launch_leduc_fsm_block<ObsT, 384>(p, tab, s);
    case 512: return launch_leduc_fsm_block<ObsT, 512>(p, tab, s);
    case 640: return launch_leduc_fsm_block<ObsT, 640>(p, tab, s);
    case 768: return launch_leduc_fsm_block<ObsT, 768>(p, tab, s);
    case 896: return launch_leduc_fsm_block<ObsT, 896>(p, tab, s);
    case 1024: return launch_leduc_fsm_block<ObsT, 1024>(p, tab, s);
    default: return launch_leduc_fsm_block<ObsT, 64>(p, tab, s);
    }
}

cudaError_t dispatch_leduc(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    // the bench / training case (throughput mode, every trajectory stream, aligned rows) runs the tabulated engine
    const uint4 *tab = nullptr;
    if (op == kOpRollout && chance == RLC_CHANCE_PHILOX && !(p.flags & kFlagNoFsm) && p.t_obs && p.t_mask && p.t_action &&
        p.t_player && p.t_done && p.t_payoffs && (tab = leduc_fsm_on_device()) != nullptr) {
        if (obs_dtype == RLC_U8 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Leduc::OBS)) & 15u) == 0)
            return launch_leduc_fsm<uint8_t>(p, tab, s);
        if (obs_dtype == RLC_F32 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Leduc::OBS * 4)) & 15u) == 0)
            return launch_leduc_fsm<float>(p, tab, s);
    }
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
