// tu_limit.cu -- kernel instantiations for Limit (one translation unit per game: parallel nvcc)
#include "game_poker.cuh"
#include "kernels.cuh"
#include <mutex>
#include <type_traits>
namespace rlc {

// ==========================================================================================
// Warp-specialised Limit Hold'em rollout (throughput mode, every trajectory stream, whole warps).
//
// 16 384 envs are 512 warps for the 592 warp schedulers of a B200: with a thread per env the kernel is bound by the
// latency of ONE warp's dependent instruction chain per env-step (obs row, flush, legal set, policy, betting transition,
// and -- episodes last ~3 steps -- a deal and now and then a showdown for some lane), at IPC 0.33.  Here every group of 32
// envs is played by FOUR warps that run concurrently (lane = env in each of them), connected by rings in shared memory:
//   ENV    the transition only: legal set -> policy (Philox) -> betting step -> payoffs -> next deal popped from the deal
//          ring; publishes a 3-word snapshot per env-step (cards, legal nibble, player, public-card count, raise counters)
//          and writes the action / done / payoffs streams;
//   EMIT   turns snapshots into trajectory rows: the 72-element obs row (tile in shared memory, 128-bit streaming stores),
//          the legal mask and the player stream;
//   DEAL   keeps the deal ring full: episode E of an env is a pure function of (seed, env, E) (deal words, common.cuh):
//          Philox block -> nine-card Fisher-Yates traceback -> 64-bit record (nine 6-bit cards + small blind);
//   JUDGE  runs the 7-card evaluator on every record behind the dealer (who wins if nobody folds), so that a showdown costs
//          the ENV warp one shared-memory read.
// Counters (deals produced / judged / released, steps published / emitted) live in shared memory, accessed volatile with
// __threadfence_block() between payload and counter; a warp only spins when it overtakes its producer.  Roles are rotated
// over the warp slots by block index so that every scheduler of an SM hosts a mix of roles.  Results are identical to
// k_rollout<Limit, ChancePhilox, ...>: same engine functions (game_poker.cuh), same deal words.
// ==========================================================================================
constexpr int kWsWarps = 4;
constexpr int kRing = 8;        // deals in flight per env
constexpr int kSnap = 4;        // env-steps ENV may run ahead of EMIT
constexpr int kPol = 2;         // policy blocks (4 env-steps each) in flight per env

__device__ __forceinline__ uint2 limit_pack_cards(const Limit &g) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) lo |= (uint32_t)g.card[k] << (6 * k);
#pragma unroll
    for (int k = 0; k < 4; k++) hi |= (uint32_t)g.card[5 + k] << (6 * k);
    return make_uint2(lo, hi);
}
__device__ __forceinline__ void limit_unpack_cards(Limit &g, uint2 d) {
#pragma unroll
    for (int k = 0; k < 5; k++) g.card[k] = (int)bf_get(d.x, 6 * k, 6);
#pragma unroll
    for (int k = 0; k < 4; k++) g.card[5 + k] = (int)bf_get(d.y, 6 * k, 6);
}
// the part of Limit::reset after the cards and the blind are known (game.py:71-103)
__device__ __forceinline__ void limit_open_episode(Limit &g, int sb) {
    g.chips0 = sb == 0 ? 1 : 2; g.chips1 = sb == 0 ? 2 : 1;
    g.fold0 = g.fold1 = 0; g.rc = 0;
    g.r.start(sb, g.chips0, g.chips1);
    g.rn_shown = g.rn;                                       // Q-LH1
    g.rn = 0;
}
struct LimitWsSmem {                 // per block (= one group of 32 envs), after the obs tile
    static constexpr int kDeal = 0;                                   // uint2 [kRing][32]        DEAL -> ENV
    static constexpr int kSnapBuf = kDeal + kRing * 32 * 8;           // uint32 [kSnap][3][32]    ENV -> EMIT
    static constexpr int kPolBuf = kSnapBuf + kSnap * 3 * 32 * 4;     // uint32 [kPol][4][32]     DEAL -> ENV (policy words)
    static constexpr int kMail = kPolBuf + kPol * 4 * 32 * 4;         // uint32 [4][32]           ENV -> JUDGE (showdown requests)
    static constexpr int kCounters = kMail + 4 * 32 * 4;              // per lane: filled, released, pol_filled, pol_released, mail_state
    static constexpr int kBytes = kCounters + 5 * 32 * 4 + 16;        // + pub, emitted, stop
};

template <class ObsT>
__global__ void __launch_bounds__(32 * kWsWarps) k_rollout_limit_ws(const KParams p) {
    extern __shared__ uint4 smem_raw[];
    constexpr int kRowBytes = Limit::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = 32 * kRowBytes;
    uint8_t *sm = reinterpret_cast<uint8_t *>(smem_raw);
    uint8_t *ws = sm + kTileBytes;
    volatile uint2 *ring = reinterpret_cast<volatile uint2 *>(ws + LimitWsSmem::kDeal);
    volatile uint32_t *snap = reinterpret_cast<volatile uint32_t *>(ws + LimitWsSmem::kSnapBuf);
    volatile uint32_t *pol = reinterpret_cast<volatile uint32_t *>(ws + LimitWsSmem::kPolBuf);
    volatile uint32_t *mail = reinterpret_cast<volatile uint32_t *>(ws + LimitWsSmem::kMail);
    volatile uint32_t *filled = reinterpret_cast<volatile uint32_t *>(ws + LimitWsSmem::kCounters);   // deals produced
    volatile uint32_t *released = filled + 32;                            // deals whose episode is over
    volatile uint32_t *pol_filled = released + 32, *pol_released = pol_filled + 32;
    volatile uint32_t *mail_state = pol_released + 32;                    // 1: a showdown waits for the JUDGE lane
    volatile uint32_t *pub = mail_state + 32, *emitted = pub + 1, *stop = pub + 2;
    const int lane = threadIdx.x & 31;
    const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) % kWsWarps);   // 0 ENV, 1 EMIT, 2 DEAL, 3 JUDGE
    const size_t i = (size_t)blockIdx.x * 32 + lane;                     // the launcher guarantees n % 32 == 0
    if (threadIdx.x < 32) {
        filled[lane] = 0; released[lane] = 0; pol_filled[lane] = 0; pol_released[lane] = 0; mail_state[lane] = 0;
        if (lane == 0) { *pub = 0; *emitted = 0; *stop = 0; }
    }
    __syncthreads();

    if (role == 2) {                                                     // ---- DEAL: policy blocks first, then deals
        const uint32_t base_ep = p.state[i];                             // episodes this env has started so far
        const uint32_t q0 = p.state[2 * p.n + i] >> 2;                   // first policy block the env will ask for (k >> 2)
        ChancePhilox ch; ch.init(p.seed, p.env_id_base + (uint32_t)i);
        uint32_t f = 0, pf = 0;
        while (*stop == 0) {
            const bool want_pol = pf - pol_released[lane] < (uint32_t)kPol;
            if (want_pol) {
                uint32_t b0, b1, b2, b3;
                philox4x32_10(q0 + pf, 0u, ch.env, (uint32_t)kDomBase, ch.k0, ch.k1, b0, b1, b2, b3);
                volatile uint32_t *dst = pol + (pf % kPol) * 128 + lane;
                dst[0] = b0; dst[32] = b1; dst[64] = b2; dst[96] = b3;
                __threadfence_block();
                pf++;
                pol_filled[lane] = pf;
            }
            const bool room = f - released[lane] < (uint32_t)kRing;
            if (room) {
                Limit g; g.rn = 0;
                ch.begin_episode(base_ep + f + 1u);
                g.reset(ch);
                const uint2 d = limit_pack_cards(g);
                const int slot = (int)(f % kRing) * 32 + lane;
                ring[slot].x = d.x | ((uint32_t)g.r.pointer << 30); ring[slot].y = d.y;
                __threadfence_block();
                f++;
                filled[lane] = f;
            }
            if (!__any_sync(0xffffffffu, room || want_pol)) __nanosleep(100);
        }
        return;
    }
    if (role == 3) {                                                     // ---- JUDGE: showdowns on request
        for (;;) {
            const uint32_t stopping = *stop;
            const bool work = mail_state[lane] != 0;
            if (work) {
                __threadfence_block();
                const volatile uint32_t *mb = mail + lane;
                uint2 d; d.x = mb[0]; d.y = mb[32];
                const uint32_t cell = mb[64], pot = mb[96];
                Limit g;
                limit_unpack_cards(g, d);
                const int oc = g.showdown_outcome();
                const float p0 = oc == 2 ? 0.f : (oc == 0 ? 0.5f : -0.5f) * (float)pot;
                st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + ((size_t)(cell / 32u) * p.n + (size_t)blockIdx.x * 32 + (cell & 31u)),
                          make_float2(p0, -p0));
                mail_state[lane] = 0;
            }
            if (!__any_sync(0xffffffffu, work)) { if (stopping) break; __nanosleep(200); }
        }
        return;
    }
    if (role == 1) {                                                     // ---- EMIT
        ObsT *tile = reinterpret_cast<ObsT *>(sm);
        ObsT *row = tile + lane * Limit::OBS;
        warp_tile_zero(reinterpret_cast<uint8_t *>(tile), kTileBytes, lane);
        __syncwarp();
        uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + (i - lane) * (size_t)kRowBytes;
        const size_t obs_step = p.n * (size_t)kRowBytes;
        size_t rowi = i;
        for (int t = 0; t < p.T; t++, rowi += p.n, o_obs += obs_step) {
            while (*pub <= (uint32_t)t) { }
            __threadfence_block();
            const volatile uint32_t *sp = snap + (t % kSnap) * 96 + lane;
            const uint32_t c_lo = sp[0], c_hi = sp[32], meta = sp[64];
            // envs/limitholdem.py:40-71 from the packed fields
            const uint32_t seat = (meta >> 4) & 1u, npub = (meta >> 5) & 7u;
            row[(c_lo >> (6 * seat)) & 63u] = (ObsT)1;
            row[(c_lo >> (12 + 6 * seat)) & 63u] = (ObsT)1;
            if (npub > 0) { row[(c_lo >> 24) & 63u] = (ObsT)1; row[c_hi & 63u] = (ObsT)1; row[(c_hi >> 6) & 63u] = (ObsT)1; }
            if (npub > 3) row[(c_hi >> 12) & 63u] = (ObsT)1;
            if (npub > 4) row[(c_hi >> 18) & 63u] = (ObsT)1;
#pragma unroll
            for (int k = 0; k < 4; k++) row[52 + 5 * k + ((meta >> (8 + 3 * k)) & 7u)] = (ObsT)1;
            __syncwarp();
            if (lane == 0) *emitted = (uint32_t)t + 1u;                  // every lane has consumed its snapshot: ENV may reuse the slot
            tile_store_begin<kTileBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
            const uint32_t m[1] = { meta & 15u };
            write_mask_row<Limit>(reinterpret_cast<uint8_t *>(p.t_mask), rowi, m);
            st_stream(p.t_player + rowi, (int)seat);
            tile_store_end<kTileBytes>(reinterpret_cast<uint8_t *>(tile), lane);
            __syncwarp();
        }
        return;
    }

    // ---- ENV
    Limit g; EnvHeader h; int err = 0;
    h.load(p.state, p.n, i);
    g.load(p.state + kHeaderWords * p.n, p.n, i);
    uint2 cards = limit_pack_cards(g);
    uint32_t popped = 0;                                                 // deals taken from the ring by this lane
    const uint32_t q0 = h.k >> 2;
    uint32_t bq = 0xffffffffu, b0 = 0, b1 = 0, b2 = 0, b3 = 0;           // policy block held in registers (index relative to q0)
    auto pop_deal = [&]() {
        released[lane] = popped;                                         // the slots of the finished episodes may be reused
        while (filled[lane] <= popped) { }
        __threadfence_block();
        const int slot = (int)(popped % kRing) * 32 + lane;
        const uint32_t x = ring[slot].x;
        cards = make_uint2(x & 0x3fffffffu, ring[slot].y);
        popped++;
        h.episode++; h.t = 0;
        limit_open_episode(g, (int)((x >> 30) & 1u));
    };
    if (h.episode == 0 || g.over()) pop_deal();
    ChancePhilox nochance; nochance.init(0, 0);
    for (int t = 0; t < p.T; t++) {
        uint32_t m[1];
        g.legal(m);
        {   // publish the snapshot of this env-step
            while ((uint32_t)t - *emitted >= (uint32_t)kSnap) { }
            const uint32_t shown = h.t == 0 ? g.rn_shown : g.rn;           // Q-LH1: the reset() state shows last episode's list
            volatile uint32_t *sp = snap + (t % kSnap) * 96 + lane;
            sp[0] = cards.x; sp[32] = cards.y;
            sp[64] = (m[0] & 15u) | ((uint32_t)g.r.pointer << 4) | ((uint32_t)g.n_public() << 5) | (shown << 8);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *pub = (uint32_t)t + 1u;
        }
        const uint32_t q = (h.k >> 2) - q0;                               // policy word W_k from the DEAL warp's blocks
        if (q != bq) {
            while (pol_filled[lane] <= q) { }
            __threadfence_block();
            const volatile uint32_t *src = pol + (q % kPol) * 128 + lane;
            b0 = src[0]; b1 = src[32]; b2 = src[64]; b3 = src[96];
        }
        const uint32_t word = sel4(b0, b1, b2, b3, h.k & 3u);
        const size_t rowi = (size_t)t * p.n + i;
        int cnt;
        const int a = pick_action<Limit>(m, word, cnt);
        st_stream(p.t_action + rowi, a);
        if (q != bq) { bq = q; pol_released[lane] = q + 1u; }            // the block is in registers (its words were just used): slot free
        g.step(a, nochance, err);
        h.t++; h.k++;
        const bool over = g.over();
        p.t_done[rowi] = over ? 1 : 0;
        if (over) {
            if (g.fold0 + g.fold1 == 1) {
                float pay[2];
                g.payoffs_given(pay, g.fold1 ? 0 : 1);
                st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, make_float2(pay[0], pay[1]));
            } else {                                                      // showdown: hand the cell to the JUDGE lane
                while (mail_state[lane] != 0) { }
                volatile uint32_t *mb = mail + lane;
                mb[0] = cards.x; mb[32] = cards.y; mb[64] = (uint32_t)t * 32u + (uint32_t)lane; mb[96] = (uint32_t)min(g.chips0, g.chips1);
                __threadfence_block();
                mail_state[lane] = 1;
            }
            pop_deal();
        } else st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, make_float2(0.f, 0.f));
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) *stop = 1;
    limit_unpack_cards(g, cards);
    h.store(p.state, p.n, i);
    g.store(p.state + kHeaderWords * p.n, p.n, i);
    if (err && p.err) p.err[i] |= err;
}

// ==========================================================================================
// Table-driven Limit Hold'em rollout (throughput mode, every trajectory stream): the default for the bench / training case.
//
// With a thread per env the kernel is bound by the number of dependent instructions one warp issues per env-step (ncu:
// 495 warp-instructions per warp-step at one warp per scheduler, IPC 0.34; the alu pipe takes one warp-instruction per two
// cycles).  Three things are taken off that chain:
//   * betting: everything about a Limit state except the cards, the chips and the raise history is one of 98 "betting
//     states" (round, pointer, raised amounts, raise / check counters, folds).  Their transition function is tabulated ONCE
//     per device by running the register engine (Limit::step / legal / over / n_public, game_poker.cuh) over every reachable
//     state -- table == engine by construction, like the Leduc automaton (tu_leduc.cu).  One uint4 per state:
//       .x  legal mask [0:4) | is_over [4] | public cards shown [5:8) | state key [11:32) (have_raised 3, not_raise_num 2,
//           pointer 1, fold0, fold1, raised0 5, raised1 5, round 3 -- the bit fields of the packed state words)
//       .y  id of the next state after the k-th legal action (ascending id order, a byte each)
//       .z  id of the k-th legal action (a byte each)      .w  chips the actor adds with it (a byte each)
//     State ids 0 / 1 open an episode with seat 0 / 1 as small blind.  Chips (one byte per seat) and the raise history
//     (4 x 3 bits) ride in two registers and are updated from the entry.
//   * deals: the deal of episode E is a pure function of (seed, env, E) (deal words, common.cuh), so every lane keeps a ring
//     of prepared deals in shared memory and the warp tops the rings up TOGETHER whenever one runs dry: the Philox block,
//     the mixed-radix decode and the nine-card Fisher-Yates traceback then run with most lanes active instead of for the
//     ~10 lanes of 32 that end an episode in any given step (the generic kernel pays a diverged deal in nearly every
//     warp-step).  A new episode is one 64-bit shared-memory read.
//   * the showdown (7-card evaluator, Limit::showdown_outcome) stays behind a branch: random play folds 96 % of the time.
// Same trajectory, state words and Philox draws as k_rollout<Limit, ChancePhilox, ObsT, 64, true, 32>.
// ==========================================================================================
constexpr int kLimFsmMax = 128;
static uint4 *g_lfsm[64];
static int g_lfsm_n[64];
static uint8_t *g_hlut[64];                  // evaluator tables (game_poker.cuh HoldemLut): card64[52] | t5[8192]
static std::mutex g_lfsm_mu;

// the evaluator's lookup tables, tabulated from the evaluator's own functions (holdem_add_card, straight_top, top_bits)
__global__ void k_holdem_build_tables(uint8_t *blob) {
    unsigned long long *card64 = reinterpret_cast<unsigned long long *>(blob);
    uint16_t *t5 = reinterpret_cast<uint16_t *>(blob + 52 * 8);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 52) {
        uint32_t m[4] = {0, 0, 0, 0};
        holdem_add_card(m, i);
        card64[i] = (unsigned long long)m[0] | ((unsigned long long)m[1] << 16) | ((unsigned long long)m[2] << 32) | ((unsigned long long)m[3] << 48);
    }
    if (i < 8192) {
        const uint32_t m = (uint32_t)i, st = (uint32_t)straight_top(m);
        uint32_t top = m;
        while (__popc(top) > 5) top &= top - 1;              // == top_bits(m, 5) wherever that is defined (<= 8 bits set)
        if (__popc(m) <= 8 && top != top_bits(m, 5)) top = 0xffffu;   // poison: the KATs would fail
        t5[i] = (uint16_t)(st ? (0x8000u | st) : top);
    }
}

__device__ __forceinline__ uint32_t lim_key(const Limit &g) {
    return (uint32_t)g.r.have_raised | ((uint32_t)g.r.not_raise_num << 3) | ((uint32_t)g.r.pointer << 5) | ((uint32_t)g.fold0 << 6) |
           ((uint32_t)g.fold1 << 7) | ((uint32_t)g.r.raised0 << 8) | ((uint32_t)g.r.raised1 << 13) | ((uint32_t)g.rc << 18);
}
__device__ __forceinline__ void lim_unkey(Limit &g, uint32_t k) {
#pragma unroll
    for (int c = 0; c < 9; c++) g.card[c] = c;
    g.chips0 = g.chips1 = 0; g.rn = g.rn_shown = 0;
    g.r.have_raised = bf_get(k, 0, 3); g.r.not_raise_num = bf_get(k, 3, 2); g.r.pointer = bf_get(k, 5, 1);
    g.fold0 = bf_get(k, 6, 1); g.fold1 = bf_get(k, 7, 1); g.r.raised0 = bf_get(k, 8, 5); g.r.raised1 = bf_get(k, 13, 5);
    g.rc = bf_get(k, 18, 3);
}
// single thread: breadth-first closure of the two opening states under Limit::step
__global__ void k_limit_build_fsm(uint4 *tab, int *count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    __shared__ uint32_t keys[kLimFsmMax];
    int n = 0;
    for (int sb = 0; sb < 2; sb++) {                         // game.py:71-103 with the blind draw fixed
        Limit g; lim_unkey(g, 0);
        g.r.start(sb, sb == 0 ? 1 : 2, sb == 0 ? 2 : 1);
        keys[n++] = lim_key(g);
    }
    ChancePhilox none; none.init(0, 0);
    for (int i = 0; i < n; i++) {
        Limit g; lim_unkey(g, keys[i]);
        uint32_t m[1]; g.legal(m);
        const bool over = g.over();
        uint32_t x = (m[0] & 15u) | (over ? 16u : 0u) | ((uint32_t)g.n_public() << 5) | (keys[i] << 11), y = 0, z = 0, w = 0;
        if (!over) {
            int kth = 0;
            for (int a = 0; a < 4; a++) {
                if (!((m[0] >> a) & 1u)) continue;
                Limit h; lim_unkey(h, keys[i]);
                int err = 0;
                const int actor = h.player();
                h.step(a, none, err);
                const int diff = actor ? h.chips1 : h.chips0;
                if (err || diff < 0 || diff > 255 || (actor ? h.chips0 : h.chips1) != 0) { *count = -2; return; }
                // the raise history moves exactly as the rollout kernel assumes: +1 in the field of the acting round on a raise
                const uint32_t rc = (uint32_t)g.rc;
                if (h.rn != (((uint32_t)g.r.have_raised + (a == kRaise ? 1u : 0u)) << (3u * rc))) { *count = -3; return; }
                const uint32_t k2 = lim_key(h);
                int j = 0;
                while (j < n && keys[j] != k2) j++;
                if (j == n) { if (n >= kLimFsmMax) { *count = -1; return; } keys[n++] = k2; }
                y |= (uint32_t)j << (8 * kth);
                z |= (uint32_t)a << (8 * kth);
                w |= (uint32_t)diff << (8 * kth);
                kth++;
            }
        }
        tab[i] = make_uint4(x, y, z, w);
    }
    *count = n;
}
// Built once per device by rlc_upload_tables(RLC_LIMIT, device, NULL, 0) (VecEnv does it at construction); until then the
// rollout runs the generic register engine (same results).
cudaError_t limit_init(int device) {
    if (device < 0 || device >= 64) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(g_lfsm_mu);
    if (g_lfsm[device]) return cudaSuccess;
    int prev = 0; cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    uint4 *tab = nullptr; int *cnt = nullptr, h = 0;
    e = cudaMalloc(&tab, sizeof(uint4) * kLimFsmMax);
    if (e == cudaSuccess) e = cudaMalloc(&cnt, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(tab, 0, sizeof(uint4) * kLimFsmMax);
    if (e == cudaSuccess) { k_limit_build_fsm<<<1, 32>>>(tab, cnt); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(&h, cnt, sizeof h, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h <= 0) e = cudaErrorUnknown;
    if (cnt) cudaFree(cnt);
    uint8_t *lut = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&lut, kHoldemLutBytes);
    if (e == cudaSuccess) { k_holdem_build_tables<<<8192 / 128, 128>>>(lut); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) { g_lfsm[device] = tab; g_lfsm_n[device] = h; g_hlut[device] = lut; }
    else { if (tab) cudaFree(tab); if (lut) cudaFree(lut); }
    cudaSetDevice(prev);
    return e;
}
// evaluator tables of the current device (built by limit_init on first use; NULL if that fails)
static const uint8_t *holdem_lut_on_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    { std::lock_guard<std::mutex> lock(g_lfsm_mu); if (g_hlut[dev]) return g_hlut[dev]; }
    if (limit_init(dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_lfsm_mu);
    return g_hlut[dev];
}
static const uint4 *limit_fsm_on_device(int &n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_lfsm_mu);
    n = g_lfsm_n[dev];
    return g_lfsm[dev];
}
__device__ __forceinline__ uint32_t lim_byte(uint32_t w, uint32_t k) { return __byte_perm(w, 0u, 0x4440u | k); }

// SPLIT = 2: the env-step chain of a group of 32 envs is cut in two and run by the two warps of a 64-thread block.  The
// automaton is so cheap (~50 instructions per step) that BOTH warps simply run it -- same table, same Philox words, same
// deals, so they stay in step without ever exchanging a value -- and each emits half of the trajectory: warp 0 the obs rows
// (tile, 128-bit streaming stores), warp 1 the mask / player / action / done / payoff streams, the showdowns and the final
// state.  The deal rings are shared: the rounds of a top-up alternate between the two warps (a round = one deal for every lane that is short),
// and one block barrier per top-up publishes them (the ring has 2 x RING slots, so a deal written in top-up r only
// replaces one that both warps consumed before top-up r - 1).  SPLIT = 1: one warp does everything (64-thread blocks
// then hold two groups).
template <class ObsT, int BLOCK, int RING, int SPLIT>
__global__ void __launch_bounds__(BLOCK) k_rollout_limit_fsm(const KParams p, const uint4 *__restrict__ gtab, int nstates, const uint4 *__restrict__ glut) {
    extern __shared__ uint4 smem_raw[];
    constexpr int kGroups = BLOCK / 32 / SPLIT, kSlots = SPLIT == 2 ? 2 * RING : RING;
    constexpr int kRowBytes = Limit::OBS * (int)sizeof(ObsT);
    constexpr int kTileBytes = kGroups * 32 * kRowBytes;
    uint4 *stab = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(smem_raw) + kTileBytes);
    uint2 *ring_all = reinterpret_cast<uint2 *>(stab + kLimFsmMax);          // [kGroups][kSlots][32]
    uint4 *slut = reinterpret_cast<uint4 *>(ring_all + kGroups * kSlots * 32);   // evaluator tables: card64[52] | t5[8192]
    for (int j = threadIdx.x; j < kLimFsmMax; j += BLOCK) stab[j] = j < nstates ? gtab[j] : make_uint4(16u, 0u, 0u, 0u);
    for (int j = threadIdx.x; j < kHoldemLutBytes / 16; j += BLOCK) slut[j] = glut[j];
    HoldemLut lut;
    lut.card64 = reinterpret_cast<const unsigned long long *>(slut);
    lut.t5 = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(slut) + 52 * 8);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = SPLIT == 2 ? wib >> 1 : wib, role = SPLIT == 2 ? wib & 1 : 0;   // role 0 emits obs rows, role 1 the other streams
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" :: "r"(1 + grp) : "memory"); };   // the two warps of a group
    const bool obs_role = SPLIT == 1 || role == 0, str_role = SPLIT == 1 || role == 1;
    const size_t warp_env0 = ((size_t)blockIdx.x * kGroups + grp) * 32;
    if (warp_env0 >= p.n) return;                                            // both warps of a group: no barrier is left waiting
    const size_t i = warp_env0 + lane;
    const bool valid = i < p.n;
    const int nvalid = (int)min((size_t)32, p.n - warp_env0);
    ObsT *tile = reinterpret_cast<ObsT *>(smem_raw) + (size_t)grp * 32 * Limit::OBS;
    ObsT *row = tile + lane * Limit::OBS;
    uint2 *ring = ring_all + (size_t)grp * kSlots * 32 + lane;               // deal of episode E: ring[(E % kSlots) * 32]
    if (obs_role) warp_tile_zero(reinterpret_cast<uint8_t *>(tile), 32 * kRowBytes, lane);
    __syncwarp();

    EnvHeader h; h.episode = 0; h.t = 0; h.k = 0;
    ChancePhilox ch; ch.init(p.seed, p.env_id_base + (uint32_t)i);
    uint32_t c_lo = 0, c_hi = 0, sid = 0;      // cards 0..4 / 5..8, six bits each (state words 0 / 1) ; betting-state id
    uint32_t chips = 0, rn = 0, rn_shown = 0;  // chips0 | chips1 << 8 ; raise counters of the four rounds, 3 bits each (Q-LH1: + the list the reset() state shows)
    bool fresh = false;                        // the env starts a new episode before its first step
    if (valid) {
        h.load(p.state, p.n, i);
        const uint32_t *gw = p.state + kHeaderWords * p.n;
        const uint32_t w0 = gw[i], w1 = gw[p.n + i], w2 = gw[2 * p.n + i], w3 = gw[3 * p.n + i];
        c_lo = w0 & 0x3fffffffu; c_hi = w1 & 0xffffffu;
        chips = bf_get(w2, 0, 6) | (bf_get(w2, 6, 6) << 8);
        rn = w3 & 0xfffu; rn_shown = (w3 >> 12) & 0xfffu;
        const uint32_t key = bf_get(w1, 24, 3) | (bf_get(w1, 27, 2) << 3) | (bf_get(w1, 29, 1) << 5) | (bf_get(w1, 30, 1) << 6) |
                             (bf_get(w1, 31, 1) << 7) | (bf_get(w2, 12, 5) << 8) | (bf_get(w2, 17, 5) << 13) | (bf_get(w2, 22, 3) << 18);
        int found = -1;
        for (int j = 0; j < nstates; j++) if ((stab[j].x >> 11) == key) { found = j; break; }
        // never dealt, a finished episode left by rlc_step without auto reset (or a state outside the table): deal first
        fresh = h.episode == 0 || found < 0 || ((stab[found].x >> 4) & 1u);
        sid = found < 0 ? 0u : (uint32_t)found;
    }
    if constexpr (SPLIT == 2) pair_sync();                    // both warps have read the state words before either may store them
    uint32_t filled = h.episode;               // deals prepared for every episode ordinal <= filled
    // top the deal rings up: every lane short of `cap` prepared deals gets one per round, all such lanes together
    auto refill = [&](uint32_t cap) {
        for (uint32_t round = 0;; round++) {
            const bool need = valid && filled - h.episode < cap;
            if (!__any_sync(0xffffffffu, need)) break;
            if (need) {
                filled++;
                if (SPLIT == 1 || (round & 1u) == (uint32_t)role) {            // whole rounds alternate between the two warps
                    Limit g; g.rn = 0;
                    ChancePhilox dc = ch;
                    dc.begin_episode(filled);
                    g.reset(dc);                               // deal words (episode-keyed) -> nine cards + small blind
                    const uint2 d = limit_pack_cards(g);
                    ring[(filled % kSlots) * 32] = make_uint2(d.x | ((uint32_t)g.r.pointer << 30), d.y);
                }
            }
        }
        if constexpr (SPLIT == 2) pair_sync(); else __syncwarp();
    };
    auto open_episode = [&]() {                // game.py:46-103 once the deal is known
        h.episode++; h.t = 0;
        const uint2 d = ring[(h.episode % kSlots) * 32];
        c_lo = d.x & 0x3fffffffu; c_hi = d.y;
        sid = d.x >> 30;                                       // states 0 / 1: seat 0 / 1 is the small blind
        chips = sid ? (2u | (1u << 8)) : (1u | (2u << 8));
        rn_shown = rn; rn = 0;                                 // Q-LH1
    };
    refill((uint32_t)min(RING, p.T + 1));
    if (fresh) open_episode();
    uint4 e = stab[sid];
    uint8_t *o_obs = reinterpret_cast<uint8_t *>(p.t_obs) + warp_env0 * (size_t)kRowBytes;
    const size_t obs_step = p.n * (size_t)kRowBytes;
    const bool full_warp = nvalid == 32;
    size_t rowi = i;
    auto run = [&](auto full_c) {
    constexpr bool kFullWarp = decltype(full_c)::value;
    const bool live = kFullWarp || valid;
    for (int t = 0; t < p.T; t++, rowi += p.n, o_obs += obs_step) {
        if (__any_sync(0xffffffffu, live && filled == h.episode))         // some lane could not open its next episode
            refill((uint32_t)min(RING, p.T - t));
        const uint32_t legal = e.x & 15u, ptr = (e.x >> 16) & 1u;
        if (obs_role) {
            if (live) {                        // envs/limitholdem.py:40-71 (Limit::encode_obs): hole cards of the acting seat,
                const uint32_t sh = 6u * ptr;  // the public cards shown in this round, the four raise counters one-hot
                row[(c_lo >> sh) & 63u] = (ObsT)1; row[(c_lo >> (sh + 12u)) & 63u] = (ObsT)1;
                const uint32_t np_ = (e.x >> 5) & 7u;
                if (np_ >= 3u) { row[(c_lo >> 24) & 63u] = (ObsT)1; row[c_hi & 63u] = (ObsT)1; row[(c_hi >> 6) & 63u] = (ObsT)1; }
                if (np_ >= 4u) row[(c_hi >> 12) & 63u] = (ObsT)1;
                if (np_ >= 5u) row[(c_hi >> 18) & 63u] = (ObsT)1;
                const uint32_t shown = h.t == 0 ? rn_shown : rn;
                row[52 + (shown & 7u)] = (ObsT)1; row[57 + ((shown >> 3) & 7u)] = (ObsT)1;
                row[62 + ((shown >> 6) & 7u)] = (ObsT)1; row[67 + ((shown >> 9) & 7u)] = (ObsT)1;
            }
            __syncwarp();
            if constexpr (kFullWarp) warp_tile_flush_full<32 * kRowBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
            else warp_tile_flush(o_obs, reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
        }
        if (live) {
            if (str_role) {
                st_stream(reinterpret_cast<uint32_t *>(p.t_mask) + rowi, (legal * 0x00204081u) & 0x01010101u);   // bit a -> byte a
                st_stream(p.t_player + rowi, (int)ptr);
            }
            const uint32_t word = ch.begin_step(h.k);
            const uint32_t kth = __umulhi(word, (uint32_t)__popc(legal));     // uniform over the legal ids, ascending
            const uint32_t a = lim_byte(e.z, kth);
            if (str_role) st_stream(p.t_action + rowi, (int)a);
            chips += lim_byte(e.w, kth) << (8u * ptr);
            rn += (a == (uint32_t)kRaise ? 1u : 0u) << (3u * (e.x >> 29));
            sid = lim_byte(e.y, kth);
            e = stab[sid];
            h.t++; h.k++;
            const bool over = (e.x >> 4) & 1u;
            float2 pay = make_float2(0.f, 0.f);
            if (over) {                        // game.py:233-243, judger.py:11-108 for two players
                if (str_role) {
                    const uint32_t f0 = (e.x >> 17) & 1u, f1 = (e.x >> 18) & 1u;
                    int oc = f1 ? 0 : 1;
                    if ((f0 | f1) == 0u)       // showdown: random play rarely gets here, whole warps skip the evaluator
                        oc = holdem_showdown_lut(c_lo, c_hi, lut);        // == Limit::showdown_outcome() (checked by rlc_judge_holdem)
                    const float pot = (float)min(chips & 255u, chips >> 8);
                    const float p0 = oc == 2 ? 0.f : (oc == 0 ? 0.5f : -0.5f) * pot;
                    pay = make_float2(p0, -p0);
                }
                open_episode();
                e = stab[sid];
            }
            if (str_role) {
                p.t_done[rowi] = over ? 1 : 0;
                st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, pay);
            }
        }
        __syncwarp();
    }
    };
    if (full_warp) run(std::true_type{}); else run(std::false_type{});
    if (valid && str_role) {
        h.store(p.state, p.n, i);
        uint32_t *gw = p.state + kHeaderWords * p.n;
        const uint32_t k = e.x >> 11;
        gw[i] = c_lo;
        gw[p.n + i] = c_hi | (bf_get(k, 0, 3) << 24) | (bf_get(k, 3, 2) << 27) | (bf_get(k, 5, 1) << 29) | (bf_get(k, 6, 1) << 30) | (bf_get(k, 7, 1) << 31);
        gw[2 * p.n + i] = (chips & 63u) | (((chips >> 8) & 63u) << 6) | (bf_get(k, 8, 5) << 12) | (bf_get(k, 13, 5) << 17) | (bf_get(k, 18, 3) << 22);
        gw[3 * p.n + i] = rn | (rn_shown << 12);
    }
}

// ==========================================================================================
// k_rollout_limit_pipe: the tabulated rollout as a PIPELINE of role warps per group of 32 envs (one block = one group).
//   ENV   (1 warp)   runs only the betting automaton + policy words and appends one 16-byte record per env-step to a
//                    double-buffered record chunk (kPipeChunk steps); nothing it does waits on the emission
//   DEAL  (ND warps) prepare the episode-keyed deals (Limit::reset, deal words) ahead of the ENV warp into per-lane rings;
//                    a ring slot carries the episode tag in its spare byte, so ENV validates a deal with the ONE 64-bit
//                    shared-memory load it needs anyway (no flag, no fence on its chain)
//   EMIT  (NE warps) turn records into trajectory rows; the steps of a chunk are dealt round-robin over the EMIT warps, so the
//                    ~200-instruction emission of one step has NE step-times to finish; they also stage the evaluator's
//                    lookup tables into shared memory while ENV plays its first chunk
// Chunks are handed over with named barriers (bar.arrive / bar.sync, ids 1..4: full[2], free[2]) -- one hand-shake per 16
// steps instead of one per step (k_rollout_limit_ws polled a ring every step and lost what it gained).  Same state words,
// Philox words, deals and trajectory as k_rollout_limit_fsm / k_rollout<Limit>: only the mapping of work to warps differs.
// ==========================================================================================
#ifndef RLC_PIPE_CHUNK
#define RLC_PIPE_CHUNK 16
#endif
#ifndef RLC_PIPE_RING
#define RLC_PIPE_RING 32
#endif
#ifndef RLC_PIPE_POLICY
#define RLC_PIPE_POLICY 0
#endif
constexpr int kPipeChunk = RLC_PIPE_CHUNK;     // env-steps per record buffer
constexpr int kPipeRing = RLC_PIPE_RING;       // deal slots per lane; >= kPipeChunk + 1 (a lane opens at most one episode per step + the first)
#ifndef RLC_PIPE_SHORT_TAIL
#define RLC_PIPE_SHORT_TAIL 0
#endif
// The launch ends with the emission of the LAST chunk, which no ENV work overlaps.  RLC_PIPE_SHORT_TAIL=1 cuts the last chunk so
// that its final NE steps (one per EMIT warp) form a chunk of their own: measured SLOWER (0.0610 vs 0.0600 ms, pass p10), off.
constexpr bool kPipeShortTail = RLC_PIPE_SHORT_TAIL != 0 && RLC_PIPE_POLICY == 0;
constexpr bool kPipePolicyByDeal = RLC_PIPE_POLICY != 0;   // policy words prepared by DEAL warp 0 (measured slower than four words per Philox block in ENV)
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t *q) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(q)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t *q, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(q)), "r"(v) : "memory");
}

template <class ObsT, int NE, int ND>
__global__ void __launch_bounds__(32 * (1 + ND + NE)) k_rollout_limit_pipe(const KParams p, const uint4 *__restrict__ gtab, int nstates,
                                                                            const uint4 *__restrict__ glut) {
    constexpr int BLOCK = 32 * (1 + ND + NE), K = kPipeChunk, RING = kPipeRing;
    constexpr int kRowBytes = Limit::OBS * (int)sizeof(ObsT), kTileBytes = 32 * kRowBytes;
    constexpr int kBarCount = 32 * (1 + NE);
    static_assert(RING >= K + 1 && ND >= 1 && ND <= 2, "see the flow-control argument below");
    extern __shared__ uint4 smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>(smem_raw);
    uint4 *stab = reinterpret_cast<uint4 *>(sm + NE * kTileBytes);
    uint4 *recs = stab + kLimFsmMax;                                         // [2][K][32]
    uint2 *ring_all = reinterpret_cast<uint2 *>(recs + 2 * K * 32);          // [RING][32]
    uint32_t *ctl = reinterpret_cast<uint32_t *>(ring_all + RING * 32);      // [0..32) episodes consumed per lane | 32: steps done | 33: finished | [64..96) h.k at launch
    uint32_t *pol = ctl + 96;                                                // [2][K][32] policy words W_k of a chunk (DEAL warp 0 -> ENV)
    uint4 *slut = reinterpret_cast<uint4 *>(pol + 2 * K * 32);               // evaluator tables, staged by the EMIT warps (only they judge showdowns)
#ifndef RLC_PIPE_ENV_LAST
#define RLC_PIPE_ENV_LAST 0
#endif
    // role index: 0 ENV, 1..ND DEAL, then EMIT.  Measured: putting the pace-setting ENV role on the LAST warp of the block (the scheduler
    // arbitrates highest warp id first) changes nothing (0.0598 vs 0.0594 ms), so roles follow the warp index
    const int lane = threadIdx.x & 31, wib = RLC_PIPE_ENV_LAST ? (int)((threadIdx.x >> 5) + 1) % (1 + ND + NE) : (int)(threadIdx.x >> 5);
    for (int j = threadIdx.x; j < kLimFsmMax; j += BLOCK) stab[j] = j < nstates ? gtab[j] : make_uint4(16u, 0u, 0u, 0u);
    for (int j = threadIdx.x; j < RING * 32; j += BLOCK) ring_all[j] = make_uint2(0u, 0u);      // tag 0 = no deal
    for (int j = 32 + threadIdx.x; j < 64; j += BLOCK) ctl[j] = 0u;           // [0..32) and [64..96) are written by the ENV warp just below
    if (wib == 0) {                                                          // what the DEAL warps start from: episode ordinal and step index per env
        const size_t i0 = (size_t)blockIdx.x * 32 + lane;
        const bool v0 = i0 < p.n;
        ctl[lane] = v0 ? p.state[i0] : 0u;
        ctl[64 + lane] = v0 ? p.state[2 * p.n + i0] : 0u;
    }
    __syncthreads();                                                         // the only block-wide barrier: the roles part here
    const size_t warp_env0 = (size_t)blockIdx.x * 32;
    if (warp_env0 >= p.n) return;
    const size_t i = warp_env0 + lane;
    const bool valid = i < p.n;
    const int nvalid = (int)min((size_t)32, p.n - warp_env0);
    const int n0 = (p.T + K - 1) / K, last_len = p.T - (n0 - 1) * K;         // chunks of K steps; the last one holds last_len <= K
    const bool split_tail = kPipeShortTail && last_len > NE;
    const int nchunks = n0 + (split_tail ? 1 : 0);
    auto chunk_bounds = [&](int c, int &t0, int &steps) {
        if (c < n0 - 1) { t0 = c * K; steps = K; }
        else if (c == n0 - 1) { t0 = c * K; steps = split_tail ? last_len - NE : last_len; }
        else { t0 = p.T - NE; steps = NE; }
    };
    uint2 *ring = ring_all + lane;                                           // deal of episode E: ring[(E % RING) * 32]

    EnvHeader h; h.episode = 0; h.t = 0; h.k = 0;
    uint32_t c_lo = 0, c_hi = 0, sid = 0, chips = 0, rn = 0, rn_shown = 0;
    bool fresh = false;
    if (wib == 0) {                                                          // ENV: the state words (as in k_rollout_limit_fsm)
        if (valid) {
            h.load(p.state, p.n, i);
            const uint32_t *gw = p.state + kHeaderWords * p.n;
            const uint32_t w0 = gw[i], w1 = gw[p.n + i], w2 = gw[2 * p.n + i], w3 = gw[3 * p.n + i];
            c_lo = w0 & 0x3fffffffu; c_hi = w1 & 0xffffffu;
            chips = bf_get(w2, 0, 6) | (bf_get(w2, 6, 6) << 8);
            rn = w3 & 0xfffu; rn_shown = (w3 >> 12) & 0xfffu;
            const uint32_t key = bf_get(w1, 24, 3) | (bf_get(w1, 27, 2) << 3) | (bf_get(w1, 29, 1) << 5) | (bf_get(w1, 30, 1) << 6) |
                                 (bf_get(w1, 31, 1) << 7) | (bf_get(w2, 12, 5) << 8) | (bf_get(w2, 17, 5) << 13) | (bf_get(w2, 22, 3) << 18);
            int found = -1;
#pragma unroll 7
            for (int j = nstates - 1; j >= 0; j--) found = (stab[j].x >> 11) == key ? j : found;   // no early exit: the loads pipeline
            fresh = h.episode == 0 || found < 0 || ((stab[found < 0 ? 0 : found].x >> 4) & 1u);
            sid = found < 0 ? 0u : (uint32_t)found;
        }
    }

    if (wib == 0) {
        // ---------------------------------------------------------------- ENV
        // The step is branch-free on its main path: the deal of the NEXT episode is fetched (volatile, tag-checked) while the
        // current step runs, an episode end is a handful of selects (the two opening table entries live in registers), the
        // policy words of four steps come from one Philox block computed ahead of them, and every shared-memory access
        // goes through a 32-bit shared address held in a register.
        ChancePhilox ch; ch.init(p.seed, p.env_id_base + (uint32_t)i);
        uint32_t sbase;
        asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem_raw)));        // opaque: never rematerialised
        const uint32_t stab_s = sbase + (uint32_t)(NE * kTileBytes);
        const uint32_t recs_s = stab_s + (uint32_t)sizeof(uint4) * kLimFsmMax + 16u * (uint32_t)lane;
        const uint32_t ring_s = stab_s + (uint32_t)sizeof(uint4) * (kLimFsmMax + 2 * K * 32) + 8u * (uint32_t)lane;
        const uint32_t pol_s = ring_s - 8u * (uint32_t)lane + (uint32_t)sizeof(uint2) * RING * 32 + 4u * (96u + (uint32_t)lane);
        auto lds_tab = [&](uint32_t id) {
            uint4 v;
            asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stab_s + 16u * id));
            return v;
        };
        auto lds_deal = [&](uint32_t ep) {         // ring slot of episode ep (volatile: the DEAL warps write it)
            uint2 v;
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(ring_s + 256u * (ep % (uint32_t)RING)) : "memory");
            return v;
        };
        const uint4 e_open0 = stab[0], e_open1 = stab[1];                  // states 0 / 1: seat 0 / 1 is the small blind
        uint32_t first = h.t == 0 ? 1u : 0u;
        auto take_deal = [&](uint2 d) {            // game.py:46-103 once the deal is known
            h.episode++; first = 1u; h.t = 0;
            c_lo = d.x & 0x3fffffffu; c_hi = d.y & 0xffffffu;
            sid = d.x >> 30;
            chips = sid ? (2u | (1u << 8)) : (1u | (2u << 8));
            rn_shown = rn; rn = 0;                                 // Q-LH1
        };
        auto deal_ready = [&](uint2 d, uint32_t ep) { return (d.y >> 24) == (0x80u | (ep & 0x7fu)); };
        if (fresh) {
            uint2 d;
            do { d = lds_deal(h.episode + 1u); } while (!deal_ready(d, h.episode + 1u));
            take_deal(d);
        }
        uint4 e = lds_tab(sid);
        uint32_t rec_s = 0, word_s = 0;
        auto step = [&](uint32_t word) {
            if constexpr (kPipePolicyByDeal)                       // W_k, prepared by DEAL warp 0 (same Philox block as ch.begin_step)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(word_s) : "memory");
            const uint2 nd = lds_deal(h.episode + 1u);             // next episode's deal: normally there long before it is needed
            const uint32_t legal = e.x & 15u, ptr = (e.x >> 16) & 1u;
            const uint32_t kth = __umulhi(word, (uint32_t)__popc(legal));     // uniform over the legal ids, ascending
            const uint32_t a = lim_byte(e.z, kth);
            const uint32_t shown = first ? rn_shown : rn;
            const uint32_t sid2 = lim_byte(e.y, kth);
            chips += lim_byte(e.w, kth) << (8u * ptr);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(rec_s), "r"(c_lo | (ptr << 30)), "r"(c_hi | (sid << 24)),
                         "r"(shown | (a << 12) | (sid2 << 16)), "r"(chips) : "memory");
            rn += (a == (uint32_t)kRaise ? 1u : 0u) << (3u * (e.x >> 29));
            const uint4 e2 = lds_tab(sid2);
            h.k++; rec_s += 512u; word_s += 128u;
            const bool over = (e2.x >> 4) & 1u;
            uint2 d = nd;
            while (over && !deal_ready(d, h.episode + 1u)) d = lds_deal(h.episode + 1u);   // rare: the DEAL warps fell behind
            const uint32_t nsid = d.x >> 30;
            // episode end = selects (h.t is only needed as "first step of the episode"; it is rebuilt for the state store)
            h.t = over ? 0u : h.t + 1u;
            first = over ? 1u : 0u;
            h.episode += over ? 1u : 0u;
            c_lo = over ? (d.x & 0x3fffffffu) : c_lo; c_hi = over ? (d.y & 0xffffffu) : c_hi;
            sid = over ? nsid : sid2;
            chips = over ? (nsid ? (2u | (1u << 8)) : (1u | (2u << 8))) : chips;
            rn_shown = over ? rn : rn_shown; rn = over ? 0u : rn;
            e.x = over ? (nsid ? e_open1.x : e_open0.x) : e2.x; e.y = over ? (nsid ? e_open1.y : e_open0.y) : e2.y;
            e.z = over ? (nsid ? e_open1.z : e_open0.z) : e2.z; e.w = over ? (nsid ? e_open1.w : e_open0.w) : e2.w;
        };
        const bool full_warp = nvalid == 32;
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1;
            int t0, steps;
            chunk_bounds(c, t0, steps);
            if (c >= 2) named_sync(3 + b, kBarCount);              // the EMIT warps are done with this buffer
            if (c > 0) {                                           // flow control for the DEAL warps, once per chunk
                __threadfence_block();
                sts_volatile_u32(ctl + lane, h.episode);
                if (lane == 0) sts_volatile_u32(ctl + 32, (uint32_t)t0);
            }
            rec_s = recs_s + (uint32_t)b * (uint32_t)(K * 512);
            word_s = pol_s + (uint32_t)b * (uint32_t)(K * 128);
            if constexpr (kPipePolicyByDeal) {
                named_sync(5 + b, 64);                             // the chunk's policy words are there
                if (full_warp) {
#pragma unroll 4
                    for (int s = 0; s < steps; s++) step(0u);
                } else {
                    for (int s = 0; s < steps; s++) {
                        if (valid) step(0u); else { rec_s += 512u; word_s += 128u; }
                    }
                }
            } else if (full_warp && (steps & 3) == 0 && __all_sync(0xffffffffu, (h.k & 3u) == 0u)) {
                for (int s = 0; s < steps; s += 4) {               // one Philox block = the policy words of four steps
                    uint32_t w0, w1, w2, w3;
                    philox4x32_10(h.k >> 2, 0u, ch.env, (uint32_t)kDomBase, ch.k0, ch.k1, w0, w1, w2, w3);
                    step(w0); step(w1); step(w2); step(w3);
                }
            } else {
                for (int s = 0; s < steps; s++) {
                    if (valid) step(ch.begin_step(h.k)); else rec_s += 512u;
                }
            }
            __threadfence_block();
            named_arrive(1 + b, kBarCount);
        }
        __syncwarp();
        if (lane == 0) sts_volatile_u32(ctl + 33, 1u);             // the DEAL warps may leave
        if (valid) {
            h.store(p.state, p.n, i);
            uint32_t *gw = p.state + kHeaderWords * p.n;
            const uint32_t k = e.x >> 11;
            gw[i] = c_lo;
            gw[p.n + i] = c_hi | (bf_get(k, 0, 3) << 24) | (bf_get(k, 3, 2) << 27) | (bf_get(k, 5, 1) << 29) | (bf_get(k, 6, 1) << 30) | (bf_get(k, 7, 1) << 31);
            gw[2 * p.n + i] = (chips & 63u) | (((chips >> 8) & 63u) << 6) | (bf_get(k, 8, 5) << 12) | (bf_get(k, 13, 5) << 17) | (bf_get(k, 18, 3) << 22);
            gw[3 * p.n + i] = rn | (rn_shown << 12);
        }
    } else if (wib <= ND) {
        // ---------------------------------------------------------------- DEAL
        // Warp d deals the episodes with ordinal = d (mod ND).  A lane is topped up to `cap` deals beyond the ordinal ENV published
        // at its last chunk boundary; within a chunk a lane opens at most kPipeChunk (+ 1 at launch) episodes and cap >= that,
        // so ENV can always be served; a slot (E mod RING) is rewritten only for E - RING <= published ordinal, i.e. consumed.
        const uint32_t d = (uint32_t)(wib - 1);
        ChancePhilox ch; ch.init(p.seed, p.env_id_base + (uint32_t)i);
        uint32_t next = lds_volatile_u32(ctl + lane) + 1u;
        if (ND == 2 && (next & 1u) != d) next++;
        const uint32_t k_launch = ctl[64 + lane];
        int pol_next = 0;                                          // warp 0 also prepares the policy words, one chunk ahead of ENV
        for (;;) {
            const uint32_t fin = lds_volatile_u32(ctl + 33);
            const uint32_t cons = lds_volatile_u32(ctl + lane), tp = lds_volatile_u32(ctl + 32);
            if (kPipePolicyByDeal && d == 0u && pol_next < nchunks && (uint32_t)pol_next <= tp / (uint32_t)K + 1u) {
                const int b = pol_next & 1, t0 = pol_next * K, steps = min(K, p.T - t0);
                if (valid)
                    for (int s = 0; s < steps; s++) pol[(b * K + s) * 32 + lane] = ch.begin_step(k_launch + (uint32_t)(t0 + s));
                __threadfence_block();
                named_arrive(5 + b, 64);
                pol_next++;
                continue;
            }
            const uint32_t cap = min((uint32_t)RING, (uint32_t)p.T - tp + 1u);
            const bool need = valid && next <= cons + cap;
            if (!__any_sync(0xffffffffu, need)) {
                if (fin) break;
                __nanosleep(128);
                continue;
            }
            if (need) {
                Limit g; g.rn = 0;
                ChancePhilox dc = ch;
                dc.begin_episode(next);
                g.reset(dc);                                       // deal words (episode-keyed) -> nine cards + small blind
                const uint2 cd = limit_pack_cards(g);
                // ONE 64-bit store (explicitly, not left to the compiler): ENV's 64-bit load sees the cards and their tag together
                asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" :: "r"(smem_u32(ring + (next % RING) * 32)),
                             "r"(cd.x | ((uint32_t)g.r.pointer << 30)), "r"(cd.y | ((0x80u | (next & 0x7fu)) << 24)) : "memory");
                next += ND;
            }
        }
    } else {
        // ---------------------------------------------------------------- EMIT
        const int j = wib - 1 - ND;
        HoldemLut lut;
        lut.card64 = reinterpret_cast<const unsigned long long *>(slut);
        lut.t5 = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(slut) + 52 * 8);
        ObsT *tile = reinterpret_cast<ObsT *>(sm + (size_t)j * kTileBytes);
        ObsT *row = tile + lane * Limit::OBS;
        warp_tile_zero(reinterpret_cast<uint8_t *>(tile), kTileBytes, lane);
        for (int q = j * 32 + lane; q < kHoldemLutBytes / 16; q += NE * 32) slut[q] = glut[q];
        named_sync(7, 32 * NE);                                    // among the EMIT warps, while ENV plays its first chunk
        uint8_t *o_obs0 = reinterpret_cast<uint8_t *>(p.t_obs) + warp_env0 * (size_t)kRowBytes;
        const size_t obs_step = p.n * (size_t)kRowBytes;
        auto run = [&](auto full_c) {
        constexpr bool kFullWarp = decltype(full_c)::value;
        const bool live = kFullWarp || valid;
        for (int c = 0; c < nchunks; c++) {
            const int b = c & 1;
            int t0, steps;
            chunk_bounds(c, t0, steps);
            named_sync(1 + b, kBarCount);                          // the chunk's records are written
            for (int s = (j + NE - t0 % NE) % NE; s < steps; s += NE) {
                const int t = t0 + s;
                const size_t rowi = (size_t)t * p.n + i;
                uint4 r = make_uint4(0u, 0u, 0u, 0u), epre = r;
                if (live) { r = recs[((size_t)b * K + s) * 32 + lane]; epre = stab[r.y >> 24]; }
                const uint32_t c_lo_ = r.x & 0x3fffffffu, ptr = r.x >> 30, c_hi_ = r.y & 0xffffffu;
                if (live) {                        // envs/limitholdem.py:40-71 (Limit::encode_obs)
                    const uint32_t sh = 6u * ptr;
                    row[(c_lo_ >> sh) & 63u] = (ObsT)1; row[(c_lo_ >> (sh + 12u)) & 63u] = (ObsT)1;
                    const uint32_t np_ = (epre.x >> 5) & 7u;
                    if (np_ >= 3u) { row[(c_lo_ >> 24) & 63u] = (ObsT)1; row[c_hi_ & 63u] = (ObsT)1; row[(c_hi_ >> 6) & 63u] = (ObsT)1; }
                    if (np_ >= 4u) row[(c_hi_ >> 12) & 63u] = (ObsT)1;
                    if (np_ >= 5u) row[(c_hi_ >> 18) & 63u] = (ObsT)1;
                    const uint32_t shown = r.z & 0xfffu;
                    row[52 + (shown & 7u)] = (ObsT)1; row[57 + ((shown >> 3) & 7u)] = (ObsT)1;
                    row[62 + ((shown >> 6) & 7u)] = (ObsT)1; row[67 + ((shown >> 9) & 7u)] = (ObsT)1;
                }
                __syncwarp();
                uint8_t *o_obs = o_obs0 + (size_t)t * obs_step;
                if constexpr (kFullWarp) warp_tile_flush_full<kTileBytes>(o_obs, reinterpret_cast<uint8_t *>(tile), lane);
                else warp_tile_flush(o_obs, reinterpret_cast<uint8_t *>(tile), nvalid * kRowBytes, lane);
                if (live) {
                    const uint32_t legal = epre.x & 15u;
                    st_stream(reinterpret_cast<uint32_t *>(p.t_mask) + rowi, (legal * 0x00204081u) & 0x01010101u);   // bit a -> byte a
                    st_stream(p.t_player + rowi, (int)ptr);
                    st_stream(p.t_action + rowi, (int)((r.z >> 12) & 3u));
                    const uint32_t ex = stab[(r.z >> 16) & 255u].x;
                    const bool over = (ex >> 4) & 1u;
                    float2 pay = make_float2(0.f, 0.f);
                    if (over) {                    // game.py:233-243, judger.py:11-108 for two players
                        const uint32_t f0 = (ex >> 17) & 1u, f1 = (ex >> 18) & 1u;
                        int oc = f1 ? 0 : 1;
                        if ((f0 | f1) == 0u) oc = holdem_showdown_lut(c_lo_, c_hi_, lut);
                        const float pot = (float)min(r.w & 255u, r.w >> 8);
                        const float p0 = oc == 2 ? 0.f : (oc == 0 ? 0.5f : -0.5f) * pot;
                        pay = make_float2(p0, -p0);
                    }
                    p.t_done[rowi] = over ? 1 : 0;
                    st_stream(reinterpret_cast<float2 *>(p.t_payoffs) + rowi, pay);
                }
                __syncwarp();
            }
            if (c + 2 < nchunks) { __threadfence_block(); named_arrive(3 + b, kBarCount); }
        }
        };
        if (nvalid == 32) run(std::true_type{}); else run(std::false_type{});
    }
}

template <class ObsT, int NE, int ND>
static cudaError_t launch_limit_pipe_as(const KParams &p, const uint4 *tab, int nstates, cudaStream_t s) {
    const size_t smem = (size_t)NE * 32 * Limit::OBS * sizeof(ObsT) + sizeof(uint4) * kLimFsmMax + sizeof(uint4) * 2 * kPipeChunk * 32 +
                        sizeof(uint2) * kPipeRing * 32 + (96 + 2 * kPipeChunk * 32) * sizeof(uint32_t) + kHoldemLutBytes;
    const uint8_t *lut = holdem_lut_on_device();
    if (!lut) return cudaErrorNotReady;
    auto k = k_rollout_limit_pipe<ObsT, NE, ND>;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)((p.n + 31) / 32), 32 * (1 + ND + NE), smem, s>>>(p, tab, nstates, reinterpret_cast<const uint4 *>(lut));
    return cudaGetLastError();
}
template <class ObsT>
static cudaError_t launch_limit_pipe(const KParams &p, const uint4 *tab, int nstates, int ne, int nd, cudaStream_t s) {
    if (nd == 1) {
        if (ne == 2) return launch_limit_pipe_as<ObsT, 2, 1>(p, tab, nstates, s);
        if (ne == 4) return launch_limit_pipe_as<ObsT, 4, 1>(p, tab, nstates, s);
        return launch_limit_pipe_as<ObsT, 3, 1>(p, tab, nstates, s);
    }
    if (ne == 2) return launch_limit_pipe_as<ObsT, 2, 2>(p, tab, nstates, s);
    if (ne == 4) return launch_limit_pipe_as<ObsT, 4, 2>(p, tab, nstates, s);
    return launch_limit_pipe_as<ObsT, 3, 2>(p, tab, nstates, s);
}

template <class ObsT, int SPLIT, int BLOCK>
static cudaError_t launch_limit_fsm_split(const KParams &p, const uint4 *tab, int nstates, cudaStream_t s) {
    // deals prepared ahead per lane: measured 8 / 16 / 32 / 48 / 64: 0.0999 / 0.0957 / 0.0932 / 0.0953 / 0.0939 ms (an env needs ~40 deals per 128 steps;
    // the first fill runs with every lane active)
    constexpr int RING = 32, kGroups = BLOCK / 32 / SPLIT, kSlots = SPLIT == 2 ? 2 * RING : RING;
    const size_t smem = (size_t)kGroups * 32 * Limit::OBS * sizeof(ObsT) + sizeof(uint4) * kLimFsmMax + (size_t)kGroups * kSlots * 32 * sizeof(uint2) + kHoldemLutBytes;
    const size_t per_block = (size_t)kGroups * 32;
    const uint8_t *lut = holdem_lut_on_device();
    if (!lut) return cudaErrorNotReady;
    auto k = k_rollout_limit_fsm<ObsT, BLOCK, RING, SPLIT>;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)((p.n + per_block - 1) / per_block), BLOCK, smem, s>>>(p, tab, nstates, reinterpret_cast<const uint4 *>(lut));
    return cudaGetLastError();
}
template <class ObsT>
static cudaError_t launch_limit_fsm(const KParams &p, const uint4 *tab, int nstates, cudaStream_t s) {
    const char *sp = getenv("RLC_LIMIT_FSM");                  // "1": one warp per group; default: the two-warp split
    // groups of 32 envs per block (two warps each): 4 once the batch gives every SM more than a block's worth, else fewer so that
    // small batches still spread (16 384 envs, blocks of 64 / 128 / 256 threads: 0.0937 / 0.0934 / 0.0918 ms); RLC_LIMIT_BLOCK overrides
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms = v > 0 ? v : 148;
    }
    const size_t groups = (p.n + 31) / 32;
    const char *bs = getenv("RLC_LIMIT_BLOCK");
    const int block = bs ? atoi(bs) : (groups >= (size_t)(3 * sms / 2) ? 256 : (groups >= (size_t)(sms / 2) ? 128 : 64));
    if (sp && sp[0] == '1') {
        if (block == 32) return launch_limit_fsm_split<ObsT, 1, 32>(p, tab, nstates, s);
        if (block == 128) return launch_limit_fsm_split<ObsT, 1, 128>(p, tab, nstates, s);
        return launch_limit_fsm_split<ObsT, 1, 64>(p, tab, nstates, s);
    }
    if (block == 128) return launch_limit_fsm_split<ObsT, 2, 128>(p, tab, nstates, s);
    if (block == 256) return launch_limit_fsm_split<ObsT, 2, 256>(p, tab, nstates, s);
    return launch_limit_fsm_split<ObsT, 2, 64>(p, tab, nstates, s);
}

template <class ObsT>
static cudaError_t launch_limit_ws(const KParams &p, cudaStream_t s) {
    constexpr int kRowBytes = Limit::OBS * (int)sizeof(ObsT);
    const size_t smem = 32 * kRowBytes + LimitWsSmem::kBytes;
    k_rollout_limit_ws<ObsT><<<(unsigned)(p.n / 32), 32 * kWsWarps, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t dispatch_limit(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    // the bench / training case (throughput mode, every trajectory stream, whole warps, aligned rows)
    // opt-in (RLC_LIMIT_WS=1): measured on B200 at 16 384 envs it is 2 % faster than the generic kernel (0.1045 vs 0.1067 ms,
    // profiles/r02_limit_ws.md) -- the ENV warp's 175-instruction chain still sets the pace -- which does not pay for four
    // spin-coupled warps per group, so the generic kernel stays the default
    const char *ws = getenv("RLC_LIMIT_WS");
    if (ws && ws[0] == '1' && op == kOpRollout && chance == RLC_CHANCE_PHILOX && !(p.flags & kFlagNoFsm) && p.n % 32 == 0 && p.t_obs &&
        p.t_mask && p.t_action && p.t_player && p.t_done && p.t_payoffs) {
        if (obs_dtype == RLC_U8 && (reinterpret_cast<uintptr_t>(p.t_obs) & 15u) == 0) return launch_limit_ws<uint8_t>(p, s);
        if (obs_dtype == RLC_F32 && (reinterpret_cast<uintptr_t>(p.t_obs) & 15u) == 0) return launch_limit_ws<float>(p, s);
    }
    // default for the same case: the tabulated engine with per-lane deal rings (RLC_LIMIT_FSM=0 keeps the generic kernel)
    const char *fsm = getenv("RLC_LIMIT_FSM");
    int nstates = 0;
    const uint4 *tab = nullptr;
    if (!(fsm && fsm[0] == '0') && op == kOpRollout && chance == RLC_CHANCE_PHILOX && !(p.flags & kFlagNoFsm) && p.T > 0 && p.t_obs &&
        p.t_mask && p.t_action && p.t_player && p.t_done && p.t_payoffs && (tab = limit_fsm_on_device(nstates)) != nullptr) {
        // default: the role-warp pipeline with three EMIT warps and one DEAL warp (0.0915 -> 0.060 ms at 16 384 envs, profiles/r02_limit_pipe.md);
        // RLC_LIMIT_PIPE="<EMIT warps><DEAL warps>" picks another shape, "0" the two-warp kernel below
        const char *pipe = getenv("RLC_LIMIT_PIPE");
        const bool fsm_named = fsm && (fsm[0] == '1' || fsm[0] == '2');   // RLC_LIMIT_FSM=1|2 names the one- / two-warp kernel explicitly
        const int ne = !pipe ? (fsm_named ? 0 : 3) : (pipe[0] >= '2' && pipe[0] <= '4' ? pipe[0] - '0' : 0), nd = !pipe ? 1 : (pipe[0] && pipe[1] == '2' ? 2 : 1);
        if (ne && obs_dtype == RLC_U8 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Limit::OBS)) & 15u) == 0)
            return launch_limit_pipe<uint8_t>(p, tab, nstates, ne, nd, s);
        if (ne && obs_dtype == RLC_F32 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Limit::OBS * 4)) & 15u) == 0)
            return launch_limit_pipe<float>(p, tab, nstates, ne, nd, s);
        if (obs_dtype == RLC_U8 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Limit::OBS)) & 15u) == 0)
            return launch_limit_fsm<uint8_t>(p, tab, nstates, s);
        if (obs_dtype == RLC_F32 && ((reinterpret_cast<uintptr_t>(p.t_obs) | (p.n * Limit::OBS * 4)) & 15u) == 0)
            return launch_limit_fsm<float>(p, tab, nstates, s);
    }
    return dispatch_game<Limit>(op, chance, obs_dtype, p, s);
}

// limitholdem/utils.py:526-569 compare_hands as a standalone operator: one thread per case
// Hands are scored by the table evaluator (tables staged in shared memory, as in the rollout) AND by the branch-free
// one: a case on which the two disagree answers 255 for every seat, so the reference's known-answer vectors pin both.
__global__ void k_judge_holdem(const uint8_t *cards, int n, int P, uint8_t *winners, const uint4 *__restrict__ glut) {
    __shared__ uint4 slut[kHoldemLutBytes / 16];
    for (int j = threadIdx.x; j < kHoldemLutBytes / 16; j += blockDim.x) slut[j] = glut[j];
    __syncthreads();
    const unsigned long long *card64 = reinterpret_cast<const unsigned long long *>(slut);
    const uint16_t *t5 = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(slut) + 52 * 8);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t str[RLC_MAX_PLAYERS], best = 0;
    bool agree = true;
    for (int p = 0; p < P; p++) {
        const uint8_t *c = cards + ((size_t)i * P + p) * 7;
        str[p] = 0;
        if (c[0] != 255) {
            const int h[7] = { c[0], c[1], c[2], c[3], c[4], c[5], c[6] };
            unsigned long long m64 = 0;
            for (int k = 0; k < 7; k++) m64 |= card64[h[k]];
            const uint32_t m[4] = { (uint32_t)m64 & 0x1fffu, (uint32_t)(m64 >> 16) & 0x1fffu, (uint32_t)(m64 >> 32) & 0x1fffu, (uint32_t)(m64 >> 48) & 0x1fffu };
            const uint32_t v = holdem_strength_lut(m, t5);
            agree = agree && v == holdem_strength7(h);
            str[p] = 1u + v;
        }
        best = max(best, str[p]);
    }
    for (int p = 0; p < P; p++) winners[(size_t)i * P + p] = !agree ? 255 : ((str[p] != 0 && str[p] == best) ? 1 : 0);
}
cudaError_t judge_holdem(const uint8_t *cards, int n, int P, uint8_t *winners, cudaStream_t s) {
    if (P < 1 || P > RLC_MAX_PLAYERS) return cudaErrorInvalidValue;
    const uint8_t *lut = holdem_lut_on_device();
    if (!lut) return cudaErrorNotReady;
    k_judge_holdem<<<(n + 127) / 128, 128, 0, s>>>(cards, n, P, winners, reinterpret_cast<const uint4 *>(lut));
    return cudaGetLastError();
}
}  // namespace rlc
