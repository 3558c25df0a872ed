// tu_dmc.cu -- the DMC actor's bookkeeping on device (rlcard/agents/dmc_agent/utils.py:97-155 `act`).
//
// The reference actor runs whole episodes with env.run(is_training=True) and appends, per position p, one
// row per decision of p: state (int8 obs p saw), action (int8 feature of the action taken,
// env.get_action_feature), target (payoff of p at the end of that episode), done (this was p's last decision
// of the episode) and episode_return (payoff when done else 0); rows are handed to the learner in chunks of
// T.  Here a window of a VecEnv rollout trajectory ([T][n] rows of obs / action / player / done / payoffs,
// auto reset) is folded into per-position row pools in HBM.  One warp per env walks its T rows; rows of an
// episode that has not ended inside the window wait in the env's "open episode" store and are emitted by
// the window in which the episode ends, so no decision is ever dropped or duplicated.
#include <cstring>
#include "common.cuh"
#include "../../include/rlcard_b200.h"

namespace rlc {

const uint64_t *doudizhu_rows_on_device(int device);     // tu_doudizhu.cu (nullptr before rlc_upload_tables)

struct DmcParams {
    const uint8_t *t_obs; const int32_t *t_action, *t_player; const uint8_t *t_done; const float *t_payoffs;
    int T, n, P, obs_stride /* bytes per obs row */, F, num_actions;
    const uint64_t *ddz_rows;                       // DouDizhu: 54-d features from the action table, else one-hot
    uint8_t *open_obs; int32_t *open_action; int8_t *open_player; int32_t *open_len; int Lmax;
    uint8_t *out_state[RLC_MAX_PLAYERS]; int8_t *out_action[RLC_MAX_PLAYERS]; float *out_target[RLC_MAX_PLAYERS];
    float *out_return[RLC_MAX_PLAYERS]; uint8_t *out_done[RLC_MAX_PLAYERS];
    int32_t *out_count; int cap; int32_t *overflow;
};

// warp copy of one obs row (obs_stride bytes, a multiple of 4): 128-bit chunks in batches of four per lane, every load
// of a batch in flight before the first store (a row is read once from the trajectory and written once to a pool)
__device__ __forceinline__ void copy_row(uint8_t *dst, const uint8_t *src, int nbytes, int lane) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 15u) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        const int nchunks = nbytes >> 4;
        for (int c0 = 0; c0 < nchunks; c0 += 128) {
            uint4 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { const int c = c0 + 32 * q + lane; if (c < nchunks) v[q] = __ldcs(s4 + c); }
#pragma unroll
            for (int q = 0; q < 4; q++) { const int c = c0 + 32 * q + lane; if (c < nchunks) __stcs(d4 + c, v[q]); }
        }
    } else {
        for (int c = lane; c < nbytes; c += 32) dst[c] = src[c];
    }
}
// env.get_action_feature: envs/doudizhu.py:136-167 (54-d thermometer of the action's cards) or the default
// one-hot over num_actions (envs/env.py:217-226)
__device__ __forceinline__ void write_feature(const DmcParams &q, int8_t *dst, int action, int lane) {
    if (q.ddz_rows) {
        const uint64_t c = q.ddz_rows[action];
        for (int e = lane; e < 54; e += 32) {
            int v;
            if (e < 52) v = (int)((c >> (4 * (e >> 2))) & 15ull) > (e & 3);
            else v = ((c >> (4 * (e - 39))) & 15ull) != 0;
            dst[e] = (int8_t)v;
        }
    } else {
        for (int e = lane; e < q.F; e += 32) dst[e] = (int8_t)(e == action ? 1 : 0);
    }
}

__global__ void __launch_bounds__(128) k_dmc_collect(const DmcParams q) {
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= q.n) return;
    const size_t S = (size_t)q.obs_stride;
    uint8_t *oobs = q.open_obs + (size_t)env * q.Lmax * S;
    int32_t *oact = q.open_action + (size_t)env * q.Lmax;
    int8_t *opl = q.open_player + (size_t)env * q.Lmax;
    int len0 = q.open_len[env];
    int t_start = 0;
    for (int t = 0; t < q.T; t++) {
        if (!q.t_done[(size_t)t * q.n + env]) continue;
        if (len0 < 0) { len0 = 0; t_start = t + 1; continue; }   // the episode whose rows overflowed the open store: dropped whole
        // an episode ends at row t: its decisions are open[0, len0) followed by window rows [t_start, t]
        const int nwin = t - t_start + 1, total = len0 + nwin;
        float pay[RLC_MAX_PLAYERS]; int cnt[RLC_MAX_PLAYERS], base[RLC_MAX_PLAYERS], seen[RLC_MAX_PLAYERS];
#pragma unroll
        for (int p = 0; p < RLC_MAX_PLAYERS; p++) {
            pay[p] = p < q.P ? q.t_payoffs[((size_t)t * q.n + env) * q.P + p] : 0.f;
            cnt[p] = 0; seen[p] = 0;
        }
        for (int r0 = 0; r0 < total; r0 += 32) {              // decisions per position
            const int r = r0 + lane;
            int pl = -1;
            if (r < total) pl = r < len0 ? (int)opl[r] : q.t_player[(size_t)(t_start + r - len0) * q.n + env];
#pragma unroll
            for (int p = 0; p < RLC_MAX_PLAYERS; p++) cnt[p] += __popc(__ballot_sync(0xffffffffu, pl == p));
        }
#pragma unroll
        for (int p = 0; p < RLC_MAX_PLAYERS; p++) {          // reserve rows in the pools
            int b = 0;
            if (lane == 0 && cnt[p] > 0) b = atomicAdd(q.out_count + p, cnt[p]);
            base[p] = __shfl_sync(0xffffffffu, b, 0);
        }
        int pl_next = 0, act_next = 0;                        // (player, action) of row r + 1 load while row r is copied
        if (total > 0) {
            const size_t w0 = (size_t)(t_start - len0) * q.n + env;
            pl_next = len0 > 0 ? (int)opl[0] : q.t_player[w0];
            act_next = len0 > 0 ? oact[0] : q.t_action[w0];
        }
        for (int r = 0; r < total; r++) {                     // emit in episode order
            const bool from_open = r < len0;
            const size_t wrow = (size_t)(t_start + r - len0) * q.n + env;
            const int pl = pl_next, act = act_next;
            if (r + 1 < total) {
                const size_t wn = (size_t)(t_start + r + 1 - len0) * q.n + env;
                pl_next = r + 1 < len0 ? (int)opl[r + 1] : q.t_player[wn];
                act_next = r + 1 < len0 ? oact[r + 1] : q.t_action[wn];
            }
            int b = 0, c = 0, s = 0; float py = 0.f;
#pragma unroll
            for (int p = 0; p < RLC_MAX_PLAYERS; p++) if (pl == p) { b = base[p]; c = cnt[p]; s = seen[p]; py = pay[p]; seen[p]++; }
            const int slot = b + s;
            if (slot >= q.cap) { if (lane == 0) *q.overflow = 1; continue; }
            const bool last = s == c - 1;
            uint8_t *st = nullptr; int8_t *ac = nullptr; float *tg = nullptr, *er = nullptr; uint8_t *dn = nullptr;
#pragma unroll
            for (int p = 0; p < RLC_MAX_PLAYERS; p++)
                if (pl == p) { st = q.out_state[p]; ac = q.out_action[p]; tg = q.out_target[p]; er = q.out_return[p]; dn = q.out_done[p]; }
            copy_row(st + (size_t)slot * S, from_open ? oobs + (size_t)r * S : q.t_obs + wrow * S, (int)S, lane);
            write_feature(q, ac + (size_t)slot * q.F, act, lane);
            if (lane == 0) { tg[slot] = py; er[slot] = last ? py : 0.f; dn[slot] = last ? 1 : 0; }
        }
        len0 = 0; t_start = t + 1;
        __syncwarp();                                         // the open store was read above and is rewritten below
    }
    // decisions of the episode still running wait in the open store
    const int tail = q.T - t_start;
    if (len0 < 0) { /* still inside the dropped episode */ }
    else if (len0 + tail > q.Lmax) { if (lane == 0) *q.overflow = 2; len0 = -1; }   // drop this episode, recover at its end
    else {
        for (int r = 0; r < tail; r++) {
            const size_t wrow = (size_t)(t_start + r) * q.n + env;
            copy_row(oobs + (size_t)(len0 + r) * S, q.t_obs + wrow * S, (int)S, lane);
            if (lane == 0) { oact[len0 + r] = q.t_action[wrow]; opl[len0 + r] = (int8_t)q.t_player[wrow]; }
        }
        len0 += tail;
    }
    if (lane == 0) q.open_len[env] = len0;
}

cudaError_t dmc_collect(const rlc_info &info, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_dmc_buffers *b, cudaStream_t s) {
    DmcParams q; memset(&q, 0, sizeof q);
    q.t_obs = reinterpret_cast<const uint8_t *>(traj->obs); q.t_action = traj->action; q.t_player = traj->player;
    q.t_done = traj->done; q.t_payoffs = traj->payoffs;
    q.T = T; q.n = n; q.P = info.num_players; q.obs_stride = info.obs_stride * (obs_dtype == RLC_F32 ? 4 : 1); q.num_actions = info.num_actions;
    q.F = info.game_id == RLC_DOUDIZHU ? 54 : info.num_actions;
    if (info.game_id == RLC_DOUDIZHU) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        q.ddz_rows = doudizhu_rows_on_device(dev);
        if (!q.ddz_rows) return cudaErrorNotReady;
    }
    q.open_obs = reinterpret_cast<uint8_t *>(b->open_obs); q.open_action = b->open_action; q.open_player = b->open_player;
    q.open_len = b->open_len; q.Lmax = b->open_capacity;
    for (int p = 0; p < info.num_players; p++) {
        q.out_state[p] = reinterpret_cast<uint8_t *>(b->out_state[p]); q.out_action[p] = b->out_action[p];
        q.out_target[p] = b->out_target[p]; q.out_return[p] = b->out_episode_return[p]; q.out_done[p] = b->out_done[p];
    }
    q.out_count = b->out_count; q.cap = b->out_capacity; q.overflow = b->overflow;
    k_dmc_collect<<<(n + 3) / 4, 128, 0, s>>>(q);
    return cudaGetLastError();
}


// ==========================================================================================
// The DMC agent's view of a state (rlcard/agents/dmc_agent/model.py:91-110 predict): the legal action ids in
// ascending order and one feature row per legal action, for all envs at once.
// ==========================================================================================
// one warp per env: mask row (dense uint8 [A] or bit-packed uint32 [W]) -> ids [max_ids] (-1 padded) + count
__global__ void __launch_bounds__(128) k_legal_ids(const void *mask, int n, int A, int W, int bitpacked, int max_ids,
                                                    int32_t *ids, int32_t *count) {
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= n) return;
    int32_t *out = ids + (size_t)env * max_ids;
    int total = 0;
    if (bitpacked) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(mask) + (size_t)env * W;
        for (int base = 0; base < W; base += 32) {
            const int wi = base + lane;
            uint32_t word = wi < W ? row[wi] : 0u;
            if (wi == W - 1 && (A & 31)) word &= (1u << (A & 31)) - 1u;
            int c = __popc(word), incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int at = total + incl - c;
            while (word) {
                const int b = __ffs(word) - 1; word &= word - 1;
                if (at < max_ids) out[at] = 32 * wi + b;
                at++;
            }
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
    } else {
        const uint8_t *row = reinterpret_cast<const uint8_t *>(mask) + (size_t)env * A;
        for (int base = 0; base < A; base += 32) {
            const int a = base + lane;
            const bool on = a < A && row[a] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, on);
            const int at = total + __popc(bal & ((1u << lane) - 1u));
            if (on && at < max_ids) out[at] = a;
            total += __popc(bal);
        }
    }
    for (int k = total + lane; k < max_ids; k += 32) out[k] = -1;
    if (lane == 0) count[env] = total;
}

__global__ void __launch_bounds__(128) k_action_features(DmcParams q, const int32_t *ids, int m, int8_t *out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= m) return;
    const int a = ids[i];
    if (a >= 0 && a < q.num_actions) write_feature(q, out + (size_t)i * q.F, a, lane);
    else for (int e = lane; e < q.F; e += 32) out[(size_t)i * q.F + e] = 0;
}

cudaError_t legal_ids(const rlc_info &info, const void *mask, int n, int max_ids, int32_t *ids, int32_t *count, cudaStream_t s) {
    k_legal_ids<<<(n + 3) / 4, 128, 0, s>>>(mask, n, info.num_actions, info.mask_words, info.mask_bitpacked, max_ids, ids, count);
    return cudaGetLastError();
}
cudaError_t action_features(const rlc_info &info, const int32_t *ids, int m, int8_t *out, cudaStream_t s) {
    DmcParams q; memset(&q, 0, sizeof q);
    q.num_actions = info.num_actions; q.F = info.game_id == RLC_DOUDIZHU ? 54 : info.num_actions;
    if (info.game_id == RLC_DOUDIZHU) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        q.ddz_rows = doudizhu_rows_on_device(dev);
        if (!q.ddz_rows) return cudaErrorNotReady;
    }
    k_action_features<<<(m + 3) / 4, 128, 0, s>>>(q, ids, m, out);
    return cudaGetLastError();
}

// ==========================================================================================
// run_rl.py's data path: Env.run(is_training=True) + reorganize (rlcard/utils/utils.py:153-179) as a per-step
// stream.  For every seat the transitions [state, action, reward, next_state, done] pair each decision of the seat
// with its next decision, the last one with the seat's terminal view and its payoff.  phase 0 runs BEFORE the env
// step (the acting seat's previous decision gets its next_state = the state it sees now; the current decision
// becomes pending), phase 1 AFTER it (finished envs close every pending decision against the terminal views).
// ==========================================================================================
struct RlParams {
    const uint8_t *obs, *mask; const int32_t *player, *action; const uint8_t *done; const float *payoffs; const uint8_t *terminal_obs;
    int n, P, S /* obs row bytes */, M /* mask row bytes */;
    uint8_t *pend_obs; int32_t *pend_action; uint8_t *pend_valid;
    uint8_t *o_state[RLC_MAX_PLAYERS], *o_next[RLC_MAX_PLAYERS], *o_mask[RLC_MAX_PLAYERS], *o_done[RLC_MAX_PLAYERS];
    int32_t *o_action[RLC_MAX_PLAYERS]; float *o_reward[RLC_MAX_PLAYERS];
    int32_t *count; int cap; int32_t *overflow;
};

__device__ __forceinline__ void rl_emit(const RlParams &q, int p, const uint8_t *state, int action, float reward,
                                         const uint8_t *next, const uint8_t *mask, bool done, int lane) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(q.count + p, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= q.cap) { if (lane == 0) *q.overflow = 1; return; }
    uint8_t *st = nullptr, *nx = nullptr, *mk = nullptr, *dn = nullptr; int32_t *ac = nullptr; float *rw = nullptr;
#pragma unroll
    for (int k = 0; k < RLC_MAX_PLAYERS; k++)
        if (k == p) { st = q.o_state[k]; nx = q.o_next[k]; mk = q.o_mask[k]; dn = q.o_done[k]; ac = q.o_action[k]; rw = q.o_reward[k]; }
    copy_row(st + (size_t)slot * q.S, state, q.S, lane);
    copy_row(nx + (size_t)slot * q.S, next, q.S, lane);
    copy_row(mk + (size_t)slot * q.M, mask, q.M, lane);
    if (lane == 0) { ac[slot] = action; rw[slot] = reward; dn[slot] = done ? 1 : 0; }
}

template <int PHASE>
__global__ void __launch_bounds__(128) k_rl_feed(const RlParams q) {
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= q.n) return;
    uint8_t *pobs = q.pend_obs + (size_t)env * q.P * q.S;
    if (PHASE == 0) {
        const int p = q.player[env], a = q.action[env];
        if (a < 0 || p < 0 || p >= q.P) return;                       // env not stepped this time
        const uint8_t *cur = q.obs + (size_t)env * q.S;
        if (q.pend_valid[env * q.P + p])
            rl_emit(q, p, pobs + (size_t)p * q.S, q.pend_action[env * q.P + p], 0.f, cur, q.mask + (size_t)env * q.M, false, lane);
        __syncwarp();
        copy_row(pobs + (size_t)p * q.S, cur, q.S, lane);
        if (lane == 0) { q.pend_action[env * q.P + p] = a; q.pend_valid[env * q.P + p] = 1; }
    } else {
        if (!q.done[env]) return;
        for (int p = 0; p < q.P; p++) {
            if (!q.pend_valid[env * q.P + p]) continue;
            rl_emit(q, p, pobs + (size_t)p * q.S, q.pend_action[env * q.P + p], q.payoffs[env * q.P + p],
                    q.terminal_obs + ((size_t)env * q.P + p) * q.S, q.mask + (size_t)env * q.M, true, lane);
            __syncwarp();
            if (lane == 0) q.pend_valid[env * q.P + p] = 0;
        }
    }
}

cudaError_t rl_feed(const rlc_info &info, int phase, const rlc_buffers *env, int obs_dtype, const int32_t *actions, int n,
                    const rlc_rl_buffers *b, cudaStream_t s) {
    RlParams q; memset(&q, 0, sizeof q);
    q.obs = reinterpret_cast<const uint8_t *>(env->obs); q.mask = reinterpret_cast<const uint8_t *>(env->mask);
    q.player = env->cur_player; q.action = actions; q.done = env->done; q.payoffs = env->payoffs;
    q.terminal_obs = reinterpret_cast<const uint8_t *>(env->terminal_obs);
    q.n = n; q.P = info.num_players; q.S = info.obs_stride * (obs_dtype == RLC_F32 ? 4 : 1);
    q.M = info.mask_bitpacked ? info.mask_words * 4 : info.num_actions;
    q.pend_obs = reinterpret_cast<uint8_t *>(b->pend_obs); q.pend_action = b->pend_action; q.pend_valid = b->pend_valid;
    for (int p = 0; p < info.num_players; p++) {
        q.o_state[p] = reinterpret_cast<uint8_t *>(b->out_state[p]); q.o_next[p] = reinterpret_cast<uint8_t *>(b->out_next_state[p]);
        q.o_mask[p] = reinterpret_cast<uint8_t *>(b->out_next_mask[p]); q.o_done[p] = b->out_done[p];
        q.o_action[p] = b->out_action[p]; q.o_reward[p] = b->out_reward[p];
    }
    q.count = b->out_count; q.cap = b->out_capacity; q.overflow = b->overflow;
    if (phase == 0) k_rl_feed<0><<<(n + 3) / 4, 128, 0, s>>>(q);
    else k_rl_feed<1><<<(n + 3) / 4, 128, 0, s>>>(q);
    return cudaGetLastError();
}


// ==========================================================================================
// reorganize over a fused rollout window (rlc_reorganize): the same transition rows as k_rl_feed, produced from the
// [T][n] trajectory of rlc_rollout_random plus its terminal-state pool.  One warp per env walks its T cells; the last
// decision of each seat waits (as a cell index of this window, or in pend_obs when it came from an earlier window).
// ==========================================================================================
struct ReorgParams {
    RlParams q;
    const uint8_t *t_obs, *t_mask; const int32_t *t_action, *t_player, *t_row; const uint8_t *t_done; const float *t_payoffs;
    const uint8_t *tm_obs, *tm_mask; int T;
};

__global__ void __launch_bounds__(128) k_reorganize(const ReorgParams r) {
    const RlParams &q = r.q;
    const int lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= q.n) return;
    uint8_t *pobs = q.pend_obs + (size_t)env * q.P * q.S;
    int pend_t[RLC_MAX_PLAYERS], pend_a[RLC_MAX_PLAYERS];          // -2 none, -1 row kept in pend_obs, >= 0 cell of this window
#pragma unroll
    for (int p = 0; p < RLC_MAX_PLAYERS; p++) {
        const bool v = p < q.P && q.pend_valid[env * q.P + p];
        pend_t[p] = v ? -1 : -2;
        pend_a[p] = v ? q.pend_action[env * q.P + p] : 0;
    }
    for (int t = 0; t < r.T; t++) {
        const size_t cell = (size_t)t * q.n + env;
        const int a = r.t_action[cell];
        if (a < 0) continue;                                          // idle cell of a forced-action replay
        const int pl = r.t_player[cell];
        const uint8_t *cur = r.t_obs + cell * q.S, *curmask = r.t_mask + cell * q.M;
#pragma unroll
        for (int p = 0; p < RLC_MAX_PLAYERS; p++) {
            if (p != pl) continue;
            if (pend_t[p] != -2) {
                const uint8_t *src = pend_t[p] < 0 ? pobs + (size_t)p * q.S : r.t_obs + ((size_t)pend_t[p] * q.n + env) * q.S;
                rl_emit(q, p, src, pend_a[p], 0.f, cur, curmask, false, lane);
            }
            pend_t[p] = t; pend_a[p] = a;
        }
        if (r.t_done[cell]) {
            const int row = r.t_row[cell];
            if (row < 0) { if (lane == 0) *q.overflow = 3; }          // the terminal pool was full: transitions lost
#pragma unroll
            for (int p = 0; p < RLC_MAX_PLAYERS; p++) {
                if (p >= q.P || pend_t[p] == -2) continue;
                if (row >= 0) {
                    const uint8_t *src = pend_t[p] < 0 ? pobs + (size_t)p * q.S : r.t_obs + ((size_t)pend_t[p] * q.n + env) * q.S;
                    rl_emit(q, p, src, pend_a[p], r.t_payoffs[cell * q.P + p], r.tm_obs + ((size_t)row * q.P + p) * q.S,
                            r.tm_mask + (size_t)row * q.M, true, lane);
                }
                pend_t[p] = -2;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < RLC_MAX_PLAYERS; p++) {                       // decisions still waiting for their next state
        if (p >= q.P) continue;
        if (pend_t[p] >= 0) copy_row(pobs + (size_t)p * q.S, r.t_obs + ((size_t)pend_t[p] * q.n + env) * q.S, q.S, lane);
        if (lane == 0) { q.pend_valid[env * q.P + p] = pend_t[p] != -2; q.pend_action[env * q.P + p] = pend_a[p]; }
    }
}

cudaError_t reorganize(const rlc_info &info, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_rl_buffers *b, cudaStream_t s) {
    ReorgParams r; memset(&r, 0, sizeof r);
    RlParams &q = r.q;
    q.n = n; q.P = info.num_players; q.S = info.obs_stride * (obs_dtype == RLC_F32 ? 4 : 1);
    q.M = info.mask_bitpacked ? info.mask_words * 4 : info.num_actions;
    q.pend_obs = reinterpret_cast<uint8_t *>(b->pend_obs); q.pend_action = b->pend_action; q.pend_valid = b->pend_valid;
    for (int p = 0; p < info.num_players; p++) {
        q.o_state[p] = reinterpret_cast<uint8_t *>(b->out_state[p]); q.o_next[p] = reinterpret_cast<uint8_t *>(b->out_next_state[p]);
        q.o_mask[p] = reinterpret_cast<uint8_t *>(b->out_next_mask[p]); q.o_done[p] = b->out_done[p];
        q.o_action[p] = b->out_action[p]; q.o_reward[p] = b->out_reward[p];
    }
    q.count = b->out_count; q.cap = b->out_capacity; q.overflow = b->overflow;
    r.t_obs = reinterpret_cast<const uint8_t *>(traj->obs); r.t_mask = reinterpret_cast<const uint8_t *>(traj->mask);
    r.t_action = traj->action; r.t_player = traj->player; r.t_row = traj->terminal_row; r.t_done = traj->done;
    r.t_payoffs = traj->payoffs; r.tm_obs = reinterpret_cast<const uint8_t *>(traj->terminal_obs);
    r.tm_mask = reinterpret_cast<const uint8_t *>(traj->terminal_mask); r.T = T;
    k_reorganize<<<(n + 3) / 4, 128, 0, s>>>(r);
    return cudaGetLastError();
}

// ==========================================================================================
// np.random.RandomState.seed(list of uint32) = MT19937 init_by_array, one generator per env, state column i of
// mt[625][n] (numpy/random/_mt19937 legacy seeding; the words come from rlcard/utils/seeding.py:33-113 on the host)
// ==========================================================================================
__global__ void k_seed_mt19937(const uint32_t *key_words, const int32_t *key_len, int n, uint32_t *mt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t N = (size_t)n;
    auto w = [&](int j) -> uint32_t & { return mt[(size_t)j * N + i]; };
    w(0) = 19650218u;
    for (int j = 1; j < 624; j++) { const uint32_t prev = w(j - 1); w(j) = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)j; }
    const int len = key_len[i];
    const uint32_t k0 = key_words[2 * i], k1 = key_words[2 * i + 1];
    int a = 1, b = 0;
    for (int k = 624 > len ? 624 : len; k; k--) {
        const uint32_t prev = w(a - 1);
        w(a) = (w(a) ^ ((prev ^ (prev >> 30)) * 1664525u)) + (b ? k1 : k0) + (uint32_t)b;
        a++; b++;
        if (a >= 624) { w(0) = w(623); a = 1; }
        if (b >= len) b = 0;
    }
    for (int k = 623; k; k--) {
        const uint32_t prev = w(a - 1);
        w(a) = (w(a) ^ ((prev ^ (prev >> 30)) * 1566083941u)) - (uint32_t)a;
        a++;
        if (a >= 624) { w(0) = w(623); a = 1; }
    }
    w(0) = 0x80000000u;
    w(624) = 624u;                                                   // index: the first draw twists
}
cudaError_t seed_mt19937(const uint32_t *key_words, const int32_t *key_len, int n, uint32_t *mt, cudaStream_t s) {
    k_seed_mt19937<<<(n + 127) / 128, 128, 0, s>>>(key_words, key_len, n, mt);
    return cudaGetLastError();
}

}  // namespace rlc
