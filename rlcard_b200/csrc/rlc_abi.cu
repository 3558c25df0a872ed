// rlc_abi.cu -- the C ABI (include/rlcard_b200.h): argument checking, per-game dispatch,
// error reporting.  No torch types, no allocation of per-env memory.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "kernels.cuh"

namespace rlc {
cudaError_t dispatch_blackjack(int, int, int, const KParams &, cudaStream_t);
cudaError_t dispatch_leduc(int, int, int, const KParams &, cudaStream_t);
cudaError_t leduc_init(int device);
cudaError_t limit_init(int device);
cudaError_t dispatch_limit(int, int, int, const KParams &, cudaStream_t);
cudaError_t judge_holdem(const uint8_t *, int, int, uint8_t *, cudaStream_t);
cudaError_t judge_leduc(const int32_t *, int, float *, cudaStream_t);
#ifdef RLC_HAVE_NOLIMIT
cudaError_t dispatch_nolimit(int, int, int, const KParams &, cudaStream_t);
#endif
#ifdef RLC_HAVE_UNO
cudaError_t dispatch_uno(int, int, int, const KParams &, cudaStream_t);
cudaError_t encode_uno(const uint8_t *, const uint8_t *, int, uint8_t *, cudaStream_t);
#endif
#ifdef RLC_HAVE_SCOUT
cudaError_t dispatch_scout(int, int, int, const KParams &, cudaStream_t);
#endif
#ifdef RLC_HAVE_DOUDIZHU
cudaError_t dispatch_doudizhu(int, int, int, const KParams &, cudaStream_t);
cudaError_t doudizhu_upload(int device, const void *blob, size_t nbytes);
cudaError_t judge_doudizhu(const uint8_t *, const int32_t *, int, uint32_t *, cudaStream_t);
#endif
#ifdef RLC_HAVE_COMPACT
cudaError_t compact_trajectory(int, const rlc_trajectory *, int, size_t, uint32_t *, cudaStream_t);
#endif
#ifdef RLC_HAVE_DMC
cudaError_t dmc_collect(const rlc_info &, const rlc_trajectory *, int, int, int, const rlc_dmc_buffers *, cudaStream_t);
cudaError_t legal_ids(const rlc_info &, const void *, int, int, int32_t *, int32_t *, cudaStream_t);
cudaError_t action_features(const rlc_info &, const int32_t *, int, int8_t *, cudaStream_t);
cudaError_t rl_feed(const rlc_info &, int, const rlc_buffers *, int, const int32_t *, int, const rlc_rl_buffers *, cudaStream_t);
cudaError_t reorganize(const rlc_info &, const rlc_trajectory *, int, int, int, const rlc_rl_buffers *, cudaStream_t);
cudaError_t seed_mt19937(const uint32_t *, const int32_t *, int, uint32_t *, cudaStream_t);
#endif
}  // namespace rlc

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}

static const rlc_info kInfo[RLC_NUM_GAMES] = {
    /* game, P, A, obs_dim[4], stride, native dtype, bitpacked, mask_words, state_words, reset draws, threads/env, state layout */
    { RLC_BLACKJACK, 1, 2, {2, 0, 0, 0}, 2, RLC_U8, 0, 1, rlc::kHeaderWords + 15, 55, 1, RLC_STATE_SOA, rlc::kHeaderWords + 4, {0, 0} },
    { RLC_LEDUC, 2, 4, {36, 36, 0, 0}, 36, RLC_U8, 0, 1, rlc::kHeaderWords + 1, 6, 1, RLC_STATE_SOA, rlc::kHeaderWords + 1, {0, 0} },
    { RLC_LIMIT, 2, 4, {72, 72, 0, 0}, 72, RLC_U8, 0, 1, rlc::kHeaderWords + 4, 52, 1, RLC_STATE_SOA, rlc::kHeaderWords + 4, {0, 0} },
    { RLC_UNO, 2, 61, {240, 240, 0, 0}, 240, RLC_U8, 0, 2, rlc::kHeaderWords + 66, 256, 1, RLC_STATE_SOA, rlc::kHeaderWords + 19, {0, 0} },
    { RLC_DOUDIZHU, 3, 27472, {790, 901, 901, 0}, 912, RLC_U8, 1, 860, rlc::kHeaderWords + 20, 54, 32, RLC_STATE_ROWS, rlc::kHeaderWords + 20, {0, 0} },
    { RLC_SCOUT, 4, 204, {688, 688, 688, 688}, 688, RLC_F32, 0, 7, rlc::kHeaderWords + 23, 90, 32, RLC_STATE_ROWS, rlc::kHeaderWords + 23, {0, 0} },
#ifdef RLC_HAVE_NOLIMIT
    { RLC_NOLIMIT, 2, 5, {54, 54, 0, 0}, 54, RLC_U8, 0, 1, rlc::kHeaderWords + 4, 53, 1, RLC_STATE_SOA, rlc::kHeaderWords + 4, {0, 0} },
#else
    { RLC_NOLIMIT, 2, 5, {54, 54, 0, 0}, 54, RLC_U8, 0, 1, 0, 53, 1, RLC_STATE_SOA, 0, {0, 0} },
#endif
};

static int dispatch(int game, int op, const rlc_buffers *b, rlc::KParams &p, void *stream) {
    cudaError_t e;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (game) {
    case RLC_BLACKJACK: e = rlc::dispatch_blackjack(op, b->chance, b->obs_dtype, p, s); break;
    case RLC_LEDUC: e = rlc::dispatch_leduc(op, b->chance, b->obs_dtype, p, s); break;
    case RLC_LIMIT: e = rlc::dispatch_limit(op, b->chance, b->obs_dtype, p, s); break;
#ifdef RLC_HAVE_NOLIMIT
    case RLC_NOLIMIT: e = rlc::dispatch_nolimit(op, b->chance, b->obs_dtype, p, s); break;
#endif
#ifdef RLC_HAVE_UNO
    case RLC_UNO: e = rlc::dispatch_uno(op, b->chance, b->obs_dtype, p, s); break;
#endif
#ifdef RLC_HAVE_SCOUT
    case RLC_SCOUT: e = rlc::dispatch_scout(op, b->chance, b->obs_dtype, p, s); break;
#endif
#ifdef RLC_HAVE_DOUDIZHU
    case RLC_DOUDIZHU: e = rlc::dispatch_doudizhu(op, b->chance, b->obs_dtype, p, s); break;
#endif
    default: return fail(RLC_ENOTIMPL, "game %d has no kernels in this build", game);
    }
    if (e == cudaErrorNotReady) return fail(RLC_ENOTABLE, "game %d needs rlc_upload_tables() on this device first", game);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? RLC_EINVAL : RLC_ECUDA, "CUDA: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLC_OK;
}

static int fill(int game, const rlc_buffers *b, int n, rlc::KParams &p) {
    if (game < 0 || game >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game);
    if (!b || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
    if (!b->state) return fail(RLC_EINVAL, "state buffer is required");
    if (b->chance == RLC_CHANCE_TAPE && (!b->tape || !b->tape_pos || b->tape_stride <= 0))
        return fail(RLC_EINVAL, "tape chance needs tape, tape_pos and tape_stride");
    if (b->chance == RLC_CHANCE_MT19937 && !b->mt) return fail(RLC_EINVAL, "mt19937 chance needs the mt buffer");
    if (b->chance < 0 || b->chance > 2) return fail(RLC_EINVAL, "bad chance kind %d", b->chance);
    if (b->obs_dtype != RLC_U8 && b->obs_dtype != RLC_F32) return fail(RLC_EINVAL, "bad obs dtype %d", b->obs_dtype);
    if (kInfo[game].obs_native_dtype == RLC_F32 && b->obs_dtype != RLC_F32)
        return fail(RLC_EINVAL, "game %d has fractional obs values: obs_dtype must be RLC_F32", game);
    memset(&p, 0, sizeof p);
    p.state = b->state; p.n = (size_t)n; p.seed = b->seed; p.env_id_base = b->env_id_base;
    p.tape = b->tape; p.tape_stride = b->tape_stride; p.tape_pos = b->tape_pos; p.mt = b->mt;
    p.obs = b->obs; p.mask = b->mask; p.cur_player = b->cur_player; p.done = b->done; p.payoffs = b->payoffs;
    p.terminal_obs = b->terminal_obs; p.err = b->err;
    if (b->legal_order) {
        if (b->legal_order_stride <= 0) return fail(RLC_EINVAL, "legal_order needs a positive legal_order_stride");
        p.order = b->legal_order; p.order_stride = b->legal_order_stride;
    }
    return RLC_OK;
}

extern "C" {

int rlc_abi_version(void) { return RLC_ABI_VERSION; }
const char *rlc_last_error(void) { return g_err; }
int64_t rlc_launch_count(void) { return (int64_t)g_launches.load(); }

int rlc_game_info(int game_id, rlc_info *out) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES || !out) return fail(RLC_EINVAL, "bad game id %d", game_id);
    *out = kInfo[game_id];
    return RLC_OK;
}

int rlc_upload_tables(int game_id, int device, const void *blob, size_t nbytes) {
#ifdef RLC_HAVE_DOUDIZHU
    if (game_id == RLC_DOUDIZHU) {
        cudaError_t e = rlc::doudizhu_upload(device, blob, nbytes);
        return e == cudaSuccess ? RLC_OK : fail(RLC_ECUDA, "table upload: %s", cudaGetErrorString(e));
    }
#endif
    if (game_id == RLC_LEDUC) {      /* no blob: the betting-state table is tabulated on the device from the engine itself */
        cudaError_t e = rlc::leduc_init(device);
        return e == cudaSuccess ? RLC_OK : fail(RLC_ECUDA, "leduc table build: %s", cudaGetErrorString(e));
    }
#ifdef RLC_HAVE_LIMIT
    if (game_id == RLC_LIMIT) {      /* likewise: the 98 betting states of Limit Hold'em */
        cudaError_t e = rlc::limit_init(device);
        return e == cudaSuccess ? RLC_OK : fail(RLC_ECUDA, "limit hold'em table build: %s", cudaGetErrorString(e));
    }
#endif
    (void)device; (void)blob; (void)nbytes;
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    return RLC_OK;   /* the other games have no uploaded tables */
}

int rlc_reset(int game_id, const rlc_buffers *b, int n, const uint8_t *reset_mask, void *stream) {
    rlc::KParams p; int rc = fill(game_id, b, n, p);
    if (rc) return rc;
    p.reset_mask = reset_mask;
    return dispatch(game_id, rlc::kOpReset, b, p, stream);
}

int rlc_step(int game_id, const rlc_buffers *b, const int32_t *actions, int n, int flags, void *stream) {
    rlc::KParams p; int rc = fill(game_id, b, n, p);
    if (rc) return rc;
    if (!actions) return fail(RLC_EINVAL, "actions is required");
    p.actions = actions; p.flags = flags;
    return dispatch(game_id, rlc::kOpStep, b, p, stream);
}

int rlc_observe(int game_id, const rlc_buffers *b, const int32_t *seat, int n, void *stream) {
    rlc::KParams p; int rc = fill(game_id, b, n, p);
    if (rc) return rc;
    p.seat = seat;
    return dispatch(game_id, rlc::kOpObserve, b, p, stream);
}

int rlc_rollout_random(int game_id, const rlc_buffers *b, const rlc_trajectory *traj, int n, int k_steps, void *stream) {
    rlc::KParams p; int rc = fill(game_id, b, n, p);
    if (rc) return rc;
    if (k_steps <= 0) return fail(RLC_EINVAL, "k_steps must be positive");
    if (traj) {
        p.t_obs = traj->obs; p.t_mask = traj->mask; p.t_action = traj->action; p.t_player = traj->player;
        p.t_done = traj->done; p.t_payoffs = traj->payoffs;
        p.t_forced = traj->forced_actions;
        if (traj->terminal_row) {
            if (!traj->terminal_count || traj->terminal_capacity <= 0)
                return fail(RLC_EINVAL, "terminal_row needs terminal_count and a positive terminal_capacity");
            p.tm_obs = traj->terminal_obs; p.tm_mask = traj->terminal_mask; p.tm_row = traj->terminal_row;
            p.tm_count = traj->terminal_count; p.tm_cap = traj->terminal_capacity;
        }
        if (p.t_forced || p.tm_row) p.flags |= rlc::kFlagNoFsm;
    }
    p.T = k_steps;
    return dispatch(game_id, rlc::kOpRollout, b, p, stream);
}

static int judged(cudaError_t e) {
    if (e == cudaErrorNotReady) return fail(RLC_ENOTABLE, "rlc_upload_tables(RLC_DOUDIZHU, ...) is needed on this device first");
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? RLC_EINVAL : RLC_ECUDA, "CUDA: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLC_OK;
}

int rlc_judge_holdem(const uint8_t *cards, int n, int num_players, uint8_t *winners, void *stream) {
    if (!cards || !winners || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
    return judged(rlc::judge_holdem(cards, n, num_players, winners, reinterpret_cast<cudaStream_t>(stream)));
}

int rlc_judge_leduc(const int32_t *cases, int n, float *payoffs, void *stream) {
    if (!cases || !payoffs || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
    return judged(rlc::judge_leduc(cases, n, payoffs, reinterpret_cast<cudaStream_t>(stream)));
}

int rlc_judge_doudizhu(const uint8_t *hands, const int32_t *targets, int n, uint32_t *mask, void *stream) {
    if (!hands || !mask || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
#ifdef RLC_HAVE_DOUDIZHU
    return judged(rlc::judge_doudizhu(hands, targets, n, mask, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "doudizhu is not in this build");
#endif
}

int rlc_encode_uno(const uint8_t *hands, const uint8_t *targets, int n, uint8_t *obs, void *stream) {
    if (!hands || !targets || !obs || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
#ifdef RLC_HAVE_UNO
    return judged(rlc::encode_uno(hands, targets, n, obs, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "uno is not in this build");
#endif
}

int rlc_dmc_collect(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_dmc_buffers *b, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!traj || !b || T <= 0 || n <= 0) return fail(RLC_EINVAL, "null buffers or empty window");
    if (!traj->obs || !traj->action || !traj->player || !traj->done || !traj->payoffs)
        return fail(RLC_EINVAL, "the DMC collector needs the obs, action, player, done and payoffs streams");
    if (!b->open_obs || !b->open_action || !b->open_player || !b->open_len || !b->out_count || !b->overflow ||
        b->open_capacity <= 0 || b->out_capacity <= 0) return fail(RLC_EINVAL, "incomplete rlc_dmc_buffers");
    for (int p = 0; p < kInfo[game_id].num_players; p++)
        if (!b->out_state[p] || !b->out_action[p] || !b->out_target[p] || !b->out_episode_return[p] || !b->out_done[p])
            return fail(RLC_EINVAL, "rlc_dmc_buffers: position %d pools missing", p);
#ifdef RLC_HAVE_DMC
    return judged(rlc::dmc_collect(kInfo[game_id], traj, obs_dtype, T, n, b, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "the DMC collector is not in this build");
#endif
}

int rlc_legal_ids(int game_id, const void *mask, int n, int max_ids, int32_t *ids, int32_t *count, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!mask || !ids || !count || n <= 0 || max_ids <= 0) return fail(RLC_EINVAL, "null buffers, n <= 0 or max_ids <= 0");
#ifdef RLC_HAVE_DMC
    return judged(rlc::legal_ids(kInfo[game_id], mask, n, max_ids, ids, count, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "not in this build");
#endif
}

int rlc_action_features(int game_id, const int32_t *ids, int m, int8_t *out, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!ids || !out || m <= 0) return fail(RLC_EINVAL, "null buffers or m <= 0");
#ifdef RLC_HAVE_DMC
    return judged(rlc::action_features(kInfo[game_id], ids, m, out, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "not in this build");
#endif
}

int rlc_rl_feed(int game_id, int phase, const rlc_buffers *env, const int32_t *actions, int n, const rlc_rl_buffers *b, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!env || !b || n <= 0 || (phase != 0 && phase != 1)) return fail(RLC_EINVAL, "null buffers, n <= 0 or bad phase");
    if (!env->obs || !env->mask || !env->cur_player) return fail(RLC_EINVAL, "rlc_rl_feed needs the env's obs, mask and cur_player");
    if (phase == 0 && !actions) return fail(RLC_EINVAL, "phase 0 needs the actions about to be applied");
    if (phase == 1 && (!env->done || !env->payoffs || !env->terminal_obs))
        return fail(RLC_EINVAL, "phase 1 needs done, payoffs and terminal_obs (step with RLC_TERMINAL_OBS)");
    if (!b->pend_obs || !b->pend_action || !b->pend_valid || !b->out_count || !b->overflow || b->out_capacity <= 0)
        return fail(RLC_EINVAL, "incomplete rlc_rl_buffers");
    for (int p = 0; p < kInfo[game_id].num_players; p++)
        if (!b->out_state[p] || !b->out_action[p] || !b->out_reward[p] || !b->out_next_state[p] || !b->out_next_mask[p] || !b->out_done[p])
            return fail(RLC_EINVAL, "rlc_rl_buffers: seat %d pools missing", p);
#ifdef RLC_HAVE_DMC
    return judged(rlc::rl_feed(kInfo[game_id], phase, env, env->obs_dtype, actions, n, b, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "the transition collector is not in this build");
#endif
}

int rlc_reorganize(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_rl_buffers *b, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!traj || !b || T <= 0 || n <= 0) return fail(RLC_EINVAL, "null buffers or empty window");
    if (!traj->obs || !traj->mask || !traj->action || !traj->player || !traj->done || !traj->payoffs ||
        !traj->terminal_obs || !traj->terminal_mask || !traj->terminal_row)
        return fail(RLC_EINVAL, "rlc_reorganize needs obs, mask, action, player, done, payoffs and the terminal_* pool of the window");
    if (!b->pend_obs || !b->pend_action || !b->pend_valid || !b->out_count || !b->overflow || b->out_capacity <= 0)
        return fail(RLC_EINVAL, "incomplete rlc_rl_buffers");
    for (int p = 0; p < kInfo[game_id].num_players; p++)
        if (!b->out_state[p] || !b->out_action[p] || !b->out_reward[p] || !b->out_next_state[p] || !b->out_next_mask[p] || !b->out_done[p])
            return fail(RLC_EINVAL, "rlc_rl_buffers: seat %d pools missing", p);
#ifdef RLC_HAVE_DMC
    return judged(rlc::reorganize(kInfo[game_id], traj, obs_dtype, T, n, b, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "the transition collector is not in this build");
#endif
}

int rlc_compact_words(int game_id) {
    switch (game_id) {
    case RLC_LEDUC: return 1;
    case RLC_LIMIT: return 3;
    case RLC_UNO: return 7;
    case RLC_DOUDIZHU: return 33;
    case RLC_SCOUT: return 20;
    default: return 0;
    }
}

int rlc_compact_trajectory(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, uint32_t *out, void *stream) {
    if (game_id < 0 || game_id >= RLC_NUM_GAMES) return fail(RLC_EINVAL, "bad game id %d", game_id);
    if (!traj || !out || T <= 0 || n <= 0) return fail(RLC_EINVAL, "null buffers or empty window");
    if (!traj->obs || !traj->mask || !traj->action || !traj->player || !traj->done || !traj->payoffs)
        return fail(RLC_EINVAL, "rlc_compact_trajectory needs every dense stream of the window");
    if (rlc_compact_words(game_id) == 0) return fail(RLC_ENOTIMPL, "game %d has no compact wire format", game_id);
#ifdef RLC_HAVE_COMPACT
    return judged(rlc::compact_trajectory(game_id, traj, obs_dtype, (size_t)T * (size_t)n, out, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "not in this build");
#endif
}

int rlc_seed_mt19937(const uint32_t *key_words, const int32_t *key_len, int n, uint32_t *mt, void *stream) {
    if (!key_words || !key_len || !mt || n <= 0) return fail(RLC_EINVAL, "null buffers or n <= 0");
#ifdef RLC_HAVE_DMC
    return judged(rlc::seed_mt19937(key_words, key_len, n, mt, reinterpret_cast<cudaStream_t>(stream)));
#else
    return fail(RLC_ENOTIMPL, "not in this build");
#endif
}

}  // extern "C"
