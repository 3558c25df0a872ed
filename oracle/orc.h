/*
 * oracle/orc.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference engines' algorithms for the hot path
 * (reset/deal -> step -> legal actions -> obs encode -> payoffs).  Every function
 * cites the reference file:line it follows (paths relative to
 * /root/reference/rlcard).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (rlcard_b200/) never does.
 *
 * Parity is PINNED: tests/test_oracle_golden.py replays tests/golden/<game>.npz
 * (recorded from the live reference by tests/golden/make_golden.py) through this
 * oracle and demands exact equality of obs, legal sets, next player, is_over and
 * payoffs, plus the reference's own known-answer tests restated in tests/.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BLACKJACK = 0, ORC_LEDUC = 1, ORC_LIMIT = 2, ORC_UNO = 3, ORC_DOUDIZHU = 4, ORC_SCOUT = 5, ORC_NOLIMIT = 6, ORC_NUM_GAMES = 7 };
enum { ORC_CHANCE_TAPE = 0, ORC_CHANCE_PHILOX = 1, ORC_CHANCE_MT = 2 };
#define ORC_MAX_PLAYERS 4

/* ---- chance source: every bounded draw of the engines goes through orc_below() ---- */
typedef struct orc_chance {
    int kind;
    /* tape (replay of recorded reference draws) */
    const uint8_t *tape; int64_t tape_len, tape_pos; int tape_err;
    /* philox (throughput mode; spec shared with the CUDA kernels, see DESIGN.md "Philox streams"):
     * k = env-steps taken so far, dom = 1 inside a step / 2 reset outside a step, R = chain register,
     * draw = index of the next fresh word */
    uint32_t key0, key1, env_id, k, dom, R, draw;
    uint32_t episode;       /* philox: ordinal (1, 2, ...) of the episode being dealt: deal words are keyed by it */
    /* numpy-legacy MT19937 (np.random.RandomState) */
    uint32_t mt[624]; int mti;
    /* optional recording of the draws made (any kind) */
    uint8_t *rec; int64_t rec_cap, rec_len;
} orc_chance;

uint32_t orc_below(orc_chance *ch, uint32_t n);          /* uniform in [0, n) */
uint32_t orc_chain(orc_chance *ch, uint32_t n);          /* philox: peeled off the step's base word; else = orc_below */
uint32_t orc_philox_begin_step(orc_chance *ch, uint32_t k, uint32_t n_legal);   /* -> index of the policy's action */
uint32_t orc_deal_below(orc_chance *ch, uint32_t j, uint32_t n);   /* philox: mulhi(deal word D_j of ch->episode, n) */
void orc_philox_begin_reset(orc_chance *ch, uint32_t k);
void orc_shuffle_u8(orc_chance *ch, uint8_t *x, int n);   /* numpy legacy list shuffle */
void orc_shuffle_tail_u8(orc_chance *ch, uint8_t *x, int n, int tail);
void orc_mt_init_by_array(orc_chance *ch, const uint32_t *key, int len);
uint32_t orc_mt_next(orc_chance *ch);
void orc_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]);
uint32_t orc_philox_word(uint32_t k0, uint32_t k1, uint32_t env_id, uint32_t c0, uint32_t c1, uint32_t dom, uint32_t word);

/* ---- per-game engines ---- */
typedef struct orc_game_vt {
    const char *name;
    int num_players, num_actions;
    int obs_dim[ORC_MAX_PLAYERS];
    size_t state_size;
    void (*create)(void *st);
    int  (*reset)(void *st, orc_chance *ch);               /* -> first player */
    int  (*step)(void *st, orc_chance *ch, int action_id); /* env.step incl. _decode_action -> next player */
    int  (*legal)(const void *st, uint8_t *mask);          /* legal ids of the state env.step/reset returned; -> count */
    int  (*obs)(const void *st, int seat, float *out);     /* seat>=0: get_state(seat); seat<0: state returned by last reset/step */
    int  (*is_over)(const void *st);
    int  (*player)(const void *st);
    void (*payoffs)(const void *st, double *out);
} orc_game_vt;

const orc_game_vt *orc_game(int game_id);

/* ---- env handle API (used from Python via ctypes) ---- */
typedef struct orc_env orc_env;
orc_env *orc_env_create(int game_id);
void orc_env_destroy(orc_env *e);
void orc_env_set_tape(orc_env *e, const uint8_t *tape, int64_t len);
void orc_env_set_philox(orc_env *e, uint64_t seed, uint32_t env_id);
void orc_env_set_mt(orc_env *e, const uint32_t *key, int len);
void orc_env_record(orc_env *e, uint8_t *buf, int64_t cap);
int64_t orc_env_recorded(const orc_env *e);
int64_t orc_env_tape_pos(const orc_env *e);
int orc_env_tape_err(const orc_env *e);
int orc_env_reset(orc_env *e);
int orc_env_step(orc_env *e, int action);
int orc_env_legal(const orc_env *e, uint8_t *mask);
int orc_env_obs(const orc_env *e, int seat, float *out);
int orc_env_is_over(const orc_env *e);
int orc_env_player(const orc_env *e);
void orc_env_payoffs(const orc_env *e, double *out);
int orc_info(int game_id, int *num_players, int *num_actions, int *obs_dim /*[4]*/);

/*
 * Throughput-mode restatement (Philox chance + Philox uniform-random policy, auto
 * reset), the CPU twin of rlc_rollout_random(): envs [env0, env0+n), T steps each,
 * trajectory arrays laid out [T][n][...] (any pointer may be NULL).  `envs` holds n
 * persistent handles (created by orc_envs_create) so successive calls continue.
 * Returns the number of env-steps done.  nthreads>1 splits envs over pthreads.
 */
typedef struct orc_envs orc_envs;
orc_envs *orc_envs_create(int game_id, int n, uint64_t seed, uint32_t env0);
void orc_envs_destroy(orc_envs *v);
int64_t orc_envs_rollout(orc_envs *v, int T, float *obs, int obs_stride, uint8_t *mask, int32_t *action,
                         int32_t *player, uint8_t *done, float *payoffs, int nthreads);
int64_t orc_envs_episodes(const orc_envs *v);

/* hold'em evaluator exposed for the known-answer tests: cards are 13*suit+rank, suit S0 H1 D2 C3,
 * rank A0 2..K12 (the reference card2index.json numbering); returns a strength that orders hands
 * exactly like limitholdem/utils.py compare_hands. */
uint32_t orc_holdem_strength7(const uint8_t cards[7]);

/* judger helpers for the known-answer tests (tests/test_reference_kats.py) */
void orc_holdem_winners(const uint8_t *cards /*[P][7], first card 255 = folded*/, int P, uint8_t *win);
int orc_doudizhu_legal_for(const uint8_t *hand15, int target_action /* <0: lead */, uint8_t *mask /*[27472]*/);
void orc_leduc_judge(int r0, int r1, int pub, int c0, int c1, int f0, int f1, double *out);
int orc_uno_encode(const uint8_t *codes, int n, int target_code, float *out /*[240]*/);

#ifdef __cplusplus
}
#endif
#endif
