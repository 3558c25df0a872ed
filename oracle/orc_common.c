/*
 * oracle/orc_common.c -- CPU ORACLE (test infrastructure, see orc.h).
 * Chance sources + env handles + the throughput-mode rollout twin.
 */
#include "orc.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ MT19937 */
/* np.random.RandomState.seed(list) == init_by_array (numpy/random/_mt19937.pyx
 * _legacy_seeding); called by rlcard/utils/seeding.py:33-41. */
static void mt_init_genrand(uint32_t *mt, uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}
void orc_mt_init_by_array(orc_chance *ch, const uint32_t *key, int len) {
    uint32_t *mt = ch->mt;
    mt_init_genrand(mt, 19650218u);
    int i = 1, j = 0, k = (624 > len ? 624 : len);
    for (; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    ch->mti = 624;
    ch->kind = ORC_CHANCE_MT;
}
uint32_t orc_mt_next(orc_chance *ch) {
    uint32_t *mt = ch->mt, y;
    if (ch->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        ch->mti = 0;
    }
    y = mt[ch->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* throughput-mode draw layout (DESIGN.md "Philox streams"; rlcard_b200/csrc/common.cuh is the CUDA twin).
 * A word is Philox(ctr = (c0, c1, env_id, dom), key = seed)[word]:
 *   base word of step k      W_k = (k >> 2, 0, env, dom 0)[k & 3]
 *   policy                   index among the n legal actions = mulhi(W_k, n); chain register R = lo32(W_k * n)
 *   chain(m)                 v = mulhi(R, m), R = lo32(R * m)
 *   fresh draw j of step k   F_j = (k, j >> 2, env, dom 1)[j & 3]
 *   reset outside a step     F_j = (k, j >> 2, env, dom 2)[j & 3]; R starts as F_0, fresh draws at j = 1
 *   deal word j of episode E D_j = (E, j >> 2, env, dom 3)[j & 3], E = 1, 2, ... the ordinal of the episode being dealt
 *                            (Limit Hold'em: the deal is a pure function of (seed, env, E)) */
uint32_t orc_philox_word(uint32_t k0, uint32_t k1, uint32_t env_id, uint32_t c0, uint32_t c1, uint32_t dom, uint32_t word) {
    uint32_t ctr[4] = { c0, c1, env_id, dom }, out[4];
    orc_philox4x32_10(ctr, k0, k1, out);
    return out[word & 3u];
}
uint32_t orc_philox_begin_step(orc_chance *ch, uint32_t k, uint32_t n_legal) {
    uint32_t w = orc_philox_word(ch->key0, ch->key1, ch->env_id, k >> 2, 0u, 0u, k & 3u);
    ch->k = k; ch->dom = 1u; ch->draw = 0u;
    ch->R = w * n_legal;
    return (uint32_t)(((uint64_t)w * n_legal) >> 32);
}
uint32_t orc_deal_below(orc_chance *ch, uint32_t j, uint32_t n) {
    uint32_t w = orc_philox_word(ch->key0, ch->key1, ch->env_id, ch->episode, j >> 2, 3u, j & 3u);
    uint32_t v = (uint32_t)(((uint64_t)w * n) >> 32);
    if (ch->rec) { if (ch->rec_len < ch->rec_cap) ch->rec[ch->rec_len] = (uint8_t)v; ch->rec_len++; }
    return v;
}
void orc_philox_begin_reset(orc_chance *ch, uint32_t k) {
    ch->k = k; ch->dom = 2u; ch->draw = 1u;
    ch->R = orc_philox_word(ch->key0, ch->key1, ch->env_id, k, 0u, 2u, 0u);
}

/* ------------------------------------------------------------------ bounded draws */
/* numpy legacy random_interval / buffered_bounded_masked_uint32: mask-and-reject on 32-bit
 * words, no draw at all when the range is a single value. */
uint32_t orc_below(orc_chance *ch, uint32_t n) {
    uint32_t v = 0;
    switch (ch->kind) {
    case ORC_CHANCE_TAPE:
        if (ch->tape_pos >= ch->tape_len) { ch->tape_err |= 1; v = 0; }
        else v = ch->tape[ch->tape_pos];
        ch->tape_pos++;
        if (v >= n) { ch->tape_err |= 2; v = n - 1; }
        break;
    case ORC_CHANCE_PHILOX:
        {
            uint32_t j = ch->draw++;
            v = (uint32_t)(((uint64_t)orc_philox_word(ch->key0, ch->key1, ch->env_id, ch->k, j >> 2, ch->dom, j & 3u) * n) >> 32);
        }
        break;
    default: {
        uint32_t max = n - 1, mask = max;
        if (max == 0) { v = 0; break; }
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        while ((v = (orc_mt_next(ch) & mask)) > max) {}
    } }
    if (ch->rec) { if (ch->rec_len < ch->rec_cap) ch->rec[ch->rec_len] = (uint8_t)v; ch->rec_len++; }
    return v;
}
uint32_t orc_chain(orc_chance *ch, uint32_t n) {
    if (ch->kind != ORC_CHANCE_PHILOX) return orc_below(ch, n);
    uint64_t p = (uint64_t)ch->R * n;
    ch->R = (uint32_t)p;
    uint32_t v = (uint32_t)(p >> 32);
    if (ch->rec) { if (ch->rec_len < ch->rec_cap) ch->rec[ch->rec_len] = (uint8_t)v; ch->rec_len++; }
    return v;
}
/* RandomState.shuffle on a list / 1-d object array: for i = n-1 .. 1: j = interval(i); swap */
void orc_shuffle_u8(orc_chance *ch, uint8_t *x, int n) {
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = orc_below(ch, (uint32_t)i + 1u);
        uint8_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
}

/* Shuffle of a deck from which only the last `tail` positions are ever dealt (deck.pop()).
 * Tape / MT19937 modes make every draw the reference made; throughput (Philox) mode, whose draw
 * layout is this project's own spec, makes only the `tail` observable swaps. */
void orc_shuffle_tail_u8(orc_chance *ch, uint8_t *x, int n, int tail) {
    int stop = (ch->kind == ORC_CHANCE_PHILOX && tail < n - 1) ? n - tail : 1;
    for (int i = n - 1; i >= stop; i--) {
        uint32_t j = orc_below(ch, (uint32_t)i + 1u);
        uint8_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
}

/* ------------------------------------------------------------------ env handles */
extern const orc_game_vt orc_vt_blackjack, orc_vt_leduc, orc_vt_limit, orc_vt_uno, orc_vt_doudizhu, orc_vt_scout, orc_vt_nolimit;
const orc_game_vt *orc_game(int g) {
    switch (g) {
    case ORC_BLACKJACK: return &orc_vt_blackjack;
    case ORC_LEDUC: return &orc_vt_leduc;
    case ORC_LIMIT: return &orc_vt_limit;
    case ORC_UNO: return &orc_vt_uno;
    case ORC_DOUDIZHU: return &orc_vt_doudizhu;
    case ORC_SCOUT: return &orc_vt_scout;
    case ORC_NOLIMIT: return &orc_vt_nolimit;
    }
    return NULL;
}
struct orc_env { const orc_game_vt *vt; void *st; orc_chance ch; uint32_t t, episode, k; };

static void env_init(orc_env *e, int game_id) {
    memset(e, 0, sizeof *e);
    e->vt = orc_game(game_id);
    e->st = calloc(1, e->vt->state_size);
    e->vt->create(e->st);
}
orc_env *orc_env_create(int game_id) {
    if (!orc_game(game_id)) return NULL;
    orc_env *e = (orc_env *)malloc(sizeof *e);
    env_init(e, game_id);
    return e;
}
void orc_env_destroy(orc_env *e) { if (e) { free(e->st); free(e); } }
void orc_env_set_tape(orc_env *e, const uint8_t *tape, int64_t len) {
    e->ch.kind = ORC_CHANCE_TAPE; e->ch.tape = tape; e->ch.tape_len = len; e->ch.tape_pos = 0; e->ch.tape_err = 0;
}
void orc_env_set_philox(orc_env *e, uint64_t seed, uint32_t env_id) {
    e->ch.kind = ORC_CHANCE_PHILOX; e->ch.key0 = (uint32_t)seed; e->ch.key1 = (uint32_t)(seed >> 32);
    e->ch.env_id = env_id; e->k = 0;
    orc_philox_begin_reset(&e->ch, 0u);
}
void orc_env_set_mt(orc_env *e, const uint32_t *key, int len) { orc_mt_init_by_array(&e->ch, key, len); }
void orc_env_record(orc_env *e, uint8_t *buf, int64_t cap) { e->ch.rec = buf; e->ch.rec_cap = cap; e->ch.rec_len = 0; }
int64_t orc_env_recorded(const orc_env *e) { return e->ch.rec_len; }
int64_t orc_env_tape_pos(const orc_env *e) { return e->ch.tape_pos; }
int orc_env_tape_err(const orc_env *e) { return e->ch.tape_err; }
int orc_env_reset(orc_env *e) {
    e->t = 0;
    e->ch.episode++;
    if (e->ch.kind == ORC_CHANCE_PHILOX) orc_philox_begin_reset(&e->ch, e->k);
    return e->vt->reset(e->st, &e->ch);
}
int orc_env_step(orc_env *e, int a) {
    if (e->ch.kind == ORC_CHANCE_PHILOX) {
        uint8_t m[32768];
        (void)orc_philox_begin_step(&e->ch, e->k, (uint32_t)e->vt->legal(e->st, m));
    }
    e->t++; e->k++;
    return e->vt->step(e->st, &e->ch, a);
}
int orc_env_legal(const orc_env *e, uint8_t *mask) { return e->vt->legal(e->st, mask); }
int orc_env_obs(const orc_env *e, int seat, float *out) { return e->vt->obs(e->st, seat, out); }
int orc_env_is_over(const orc_env *e) { return e->vt->is_over(e->st); }
int orc_env_player(const orc_env *e) { return e->vt->player(e->st); }
void orc_env_payoffs(const orc_env *e, double *out) { e->vt->payoffs(e->st, out); }
int orc_info(int g, int *np_, int *na, int *od) {
    const orc_game_vt *vt = orc_game(g);
    if (!vt) return -1;
    *np_ = vt->num_players; *na = vt->num_actions;
    for (int i = 0; i < ORC_MAX_PLAYERS; i++) od[i] = vt->obs_dim[i];
    return 0;
}

/* ------------------------------------------------------------------ throughput twin */
struct orc_envs { int game_id, n; orc_env *e; int64_t episodes; uint8_t *started; };
orc_envs *orc_envs_create(int game_id, int n, uint64_t seed, uint32_t env0) {
    if (!orc_game(game_id)) return NULL;
    orc_envs *v = (orc_envs *)calloc(1, sizeof *v);
    v->game_id = game_id; v->n = n;
    v->e = (orc_env *)calloc((size_t)n, sizeof(orc_env));
    v->started = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < n; i++) { env_init(&v->e[i], game_id); orc_env_set_philox(&v->e[i], seed, env0 + (uint32_t)i); }
    return v;
}
void orc_envs_destroy(orc_envs *v) {
    if (!v) return;
    for (int i = 0; i < v->n; i++) free(v->e[i].st);
    free(v->e); free(v->started); free(v);
}
int64_t orc_envs_episodes(const orc_envs *v) { return v->episodes; }

typedef struct {
    orc_envs *v; int lo, hi, T; float *obs; int obs_stride; uint8_t *mask; int32_t *action, *player;
    uint8_t *done; float *payoffs; int64_t episodes;
} rollout_job;

/* first deal: a reset outside any step (Philox domain 2); later deals happen inside the step that ended
 * the previous episode and continue that step's chance draws */
static void env_new_episode(orc_env *e, int first) {
    if (first) { orc_philox_begin_reset(&e->ch, e->k); e->episode = 0; }
    else e->episode++;
    e->ch.episode = e->episode + 1u;
    e->t = 0;
    e->vt->reset(e->st, &e->ch);
}
static void *rollout_worker(void *arg) {
    rollout_job *J = (rollout_job *)arg;
    orc_envs *v = J->v;
    const orc_game_vt *vt = v->e[0].vt;
    int A = vt->num_actions, P = vt->num_players, n = v->n;
    uint8_t *m = (uint8_t *)malloc((size_t)A);
    float *ob = (float *)malloc(sizeof(float) * 1024);
    for (int i = J->lo; i < J->hi; i++) {
        orc_env *e = &v->e[i];
        if (!v->started[i]) { env_new_episode(e, 1); v->started[i] = 1; }
        for (int t = 0; t < J->T; t++) {
            size_t row = (size_t)t * n + i;
            int cnt = vt->legal(e->st, m);
            int pl = vt->player(e->st);
            if (J->obs) {
                int d = vt->obs(e->st, -1, ob);
                float *dst = J->obs + row * J->obs_stride;
                memcpy(dst, ob, sizeof(float) * d);
                for (int k = d; k < J->obs_stride; k++) dst[k] = 0.f;
            }
            if (J->mask) memcpy(J->mask + row * A, m, (size_t)A);
            if (J->player) J->player[row] = pl;
            /* uniform-random legal action: k-th legal id in ascending order, k from the policy stream */
            int k = (int)orc_philox_begin_step(&e->ch, e->k, (uint32_t)cnt), a = 0;   /* this step's Philox coordinates */
            for (a = 0; a < A; a++) if (m[a] && k-- == 0) break;
            if (J->action) J->action[row] = a;
            e->t++; e->k++;
            vt->step(e->st, &e->ch, a);
            int over = vt->is_over(e->st);
            if (J->done) J->done[row] = (uint8_t)over;
            if (J->payoffs) {
                double pay[ORC_MAX_PLAYERS] = {0, 0, 0, 0};
                if (over) vt->payoffs(e->st, pay);
                for (int p = 0; p < P; p++) J->payoffs[row * P + p] = (float)pay[p];
            }
            if (over) { J->episodes++; env_new_episode(e, 0); }
        }
    }
    free(m); free(ob);
    return NULL;
}
int64_t orc_envs_rollout(orc_envs *v, int T, float *obs, int obs_stride, uint8_t *mask, int32_t *action,
                         int32_t *player, uint8_t *done, float *payoffs, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > v->n) nthreads = v->n;
    rollout_job *jobs = (rollout_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    for (int k = 0; k < nthreads; k++) {
        rollout_job j = { v, (int)((int64_t)v->n * k / nthreads), (int)((int64_t)v->n * (k + 1) / nthreads), T,
                          obs, obs_stride, mask, action, player, done, payoffs, 0 };
        jobs[k] = j;
        if (nthreads > 1) pthread_create(&th[k], NULL, rollout_worker, &jobs[k]);
        else rollout_worker(&jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) { if (nthreads > 1) pthread_join(th[k], NULL); v->episodes += jobs[k].episodes; }
    free(jobs); free(th);
    return (int64_t)v->n * T;
}
