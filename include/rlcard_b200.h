/*
 * rlcard_b200.h -- C ABI of the B200-native batched card-game simulator.
 *
 * This is the drop-in boundary for the reference's Env hot path
 * (rlcard/envs/env.py:52-231 reset/step/get_state/get_payoffs/run and the
 * per-game envs + engines under rlcard/envs/<game>.py, rlcard/games/<game>/).
 * Plain pointers and sizes only; all per-env memory is owned by the caller
 * (torch tensors on the Python side) and lives in HBM.  Every entry point
 * returns 0 or a negative rlc_status; rlc_last_error() describes the failure.
 * All launches are asynchronous on the caller's stream (a cudaStream_t passed
 * as void*; NULL = legacy default stream).
 *
 * Layouts (n = number of envs of this call, P players, A actions):
 *   state     opaque packed env state, rlc_game_info.state_words uint32 words per env.  Its layout depends on
 *             rlc_info.threads_per_env: 1 (thread-per-env games) -> struct-of-arrays uint32 [state_words][n];
 *             32 (warp-per-env games: DouDizhu, Scout) -> one row per env, uint32 [n][state_words].
 *             rlc_info.state_layout says which (RLC_STATE_SOA / RLC_STATE_ROWS); slice or shard it accordingly.
 *   obs       obs_t  [n][obs_stride]    row = the reference's exact obs layout of the seat's
 *                                       _extract_state, zero padded to obs_stride (max seat dim)
 *   mask      uint8  [n][A]             1 = legal (games with A <= 256)
 *             uint32 [n][mask_words]    bit a%32 of word a/32 (DouDizhu, A = 27472)
 *   cur_player int32 [n], done uint8 [n], payoffs float [n][P], err int32 [n]
 *   rollout trajectories carry a leading [T] dimension over the same rows.
 */
#ifndef RLCARD_B200_H
#define RLCARD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLC_ABI_VERSION 2
#define RLC_MAX_PLAYERS 4

/* env ids of rlcard/envs/__init__.py:6-54 covered by this library */
enum rlc_game {
    RLC_BLACKJACK = 0,   /* 'blackjack'     */
    RLC_LEDUC = 1,       /* 'leduc-holdem'  */
    RLC_LIMIT = 2,       /* 'limit-holdem'  */
    RLC_UNO = 3,         /* 'uno'           */
    RLC_DOUDIZHU = 4,    /* 'doudizhu'      */
    RLC_SCOUT = 5,       /* 'scout'         */
    RLC_NOLIMIT = 6,     /* 'no-limit-holdem' (2 players, 100 chips each: the reference defaults) */
    RLC_NUM_GAMES = 7
};

enum rlc_status {
    RLC_OK = 0, RLC_EINVAL = -1, RLC_ENOTIMPL = -2, RLC_ECUDA = -3, RLC_ENOTABLE = -4
};

/* where the engines' bounded random draws come from (replaces game.np_random, env.py:228-231) */
enum rlc_chance {
    RLC_CHANCE_PHILOX = 0,   /* throughput: Philox4x32-10 keyed (seed, global env id, env-step index) */
    RLC_CHANCE_TAPE = 1,     /* replay: recorded np.random draws, uint8 tape per env          */
    RLC_CHANCE_MT19937 = 2   /* replay from a seed: np.random.RandomState on device           */
};

enum rlc_dtype { RLC_U8 = 0, RLC_F32 = 1 };
enum rlc_state_layout { RLC_STATE_SOA = 0, RLC_STATE_ROWS = 1 };

enum rlc_step_flags {
    RLC_AUTO_RESET = 1,      /* deal the next episode inside the step that ends one          */
    RLC_TERMINAL_OBS = 2     /* also write every seat's terminal get_state (env.py:161-164)  */
};

typedef struct rlc_info {
    int32_t game_id, num_players, num_actions;
    int32_t obs_dim[RLC_MAX_PLAYERS];   /* state_shape per seat (e.g. doudizhu 790/901/901) */
    int32_t obs_stride;                 /* elements per obs row                              */
    int32_t obs_native_dtype;           /* rlc_dtype the reference values fit losslessly      */
    int32_t mask_bitpacked;             /* 0: uint8 [n][A]; 1: uint32 [n][mask_words]         */
    int32_t mask_words;
    int32_t state_words;                /* uint32 words per env                               */
    int32_t max_tape_draws_reset;       /* draws a reset can consume (tape sizing aid)        */
    int32_t threads_per_env;            /* 1 (thread per env) or 32 (warp per env)            */
    int32_t state_layout;               /* RLC_STATE_SOA [state_words][n] or RLC_STATE_ROWS [n][state_words] */
    int32_t state_words_philox;         /* words the throughput (Philox) kernels actually read and write per env
                                           (UNO keeps multiset piles, Blackjack a deck mask: fewer than state_words) */
    int32_t reserved[2];
} rlc_info;

typedef struct rlc_buffers {
    uint32_t *state;          /* [state_words][n] or [n][state_words], see rlc_info.state_layout */
    /* chance source */
    int32_t chance;           /* rlc_chance                                                   */
    uint64_t seed;            /* PHILOX key                                                   */
    uint32_t env_id_base;     /* PHILOX: global id of env 0 of this buffer (multi-GPU shards) */
    const uint8_t *tape;      /* TAPE: [n][tape_stride] draw outcomes                         */
    int32_t tape_stride;
    int32_t *tape_pos;        /* TAPE: [n] cursor (in/out)                                    */
    uint32_t *mt;             /* MT19937: [625][n] (624 state words + index), in/out          */
    /* outputs */
    void *obs;                /* [n][obs_stride] of obs_dtype                                 */
    int32_t obs_dtype;        /* rlc_dtype                                                    */
    void *mask;               /* see header comment                                           */
    int32_t *cur_player;      /* [n]                                                          */
    uint8_t *done;            /* [n]                                                          */
    float *payoffs;           /* [n][P]                                                       */
    void *terminal_obs;       /* [n][P][obs_stride], written for done envs when RLC_TERMINAL_OBS */
    int32_t *err;             /* [n] sticky bit flags: 1 tape exhausted, 2 tape value out of range,
                                 4 illegal action replaced by the reference's fallback,
                                 8 terminal-state pool of the fused rollout was full            */
    /* ABI 2, optional: the legal ids of the state that obs/mask describe, in the INSERTION ORDER of the reference's
     * state['legal_actions'] OrderedDict (what `np.random.choice(list(legal_actions.keys()))` indexes,
     * agents/random_agent.py:27): UNO = order of first occurrence in the hand list (envs/uno.py:47-50 over
     * games/uno/round.py:96-135; replay chance modes only -- the throughput mode keeps hands as multisets and
     * lists ascending), Scout = multi-card plays, then single cards, then scouts (games/scout/round.py:225-260,
     * utils/utils.py:70-100), every other game ascending (DouDizhu's lead order is Python set order in the
     * reference, i.e. undefined).  Rows are -1 padded; written by rlc_reset / rlc_step / rlc_observe. */
    int32_t *legal_order;     /* [n][legal_order_stride]                                      */
    int32_t legal_order_stride;
} rlc_buffers;

/* trajectory buffers of rlc_rollout_random: [T][n][...], any pointer may be NULL */
typedef struct rlc_trajectory {
    void *obs;                /* [T][n][obs_stride] obs the acting player saw (obs_dtype of rlc_buffers) */
    void *mask;               /* [T][n][A] or [T][n][mask_words]                              */
    int32_t *action;          /* [T][n] action id taken                                       */
    int32_t *player;          /* [T][n] acting player                                         */
    uint8_t *done;            /* [T][n] episode ended with this action                        */
    float *payoffs;           /* [T][n][P] payoffs when done else 0                           */
    /* ---- ABI 2 (all optional, NULL / 0 = off) ----
     * forced_actions: replay of recorded action sequences through the fused loop (the action half of the replay
     * mode: pair it with RLC_CHANCE_TAPE / RLC_CHANCE_MT19937 for the chance half).  Cell [t][i] is applied with
     * Env.step semantics (illegal ids take the env's fallback and raise err flag 4); a negative id leaves env i
     * untouched for that cell (its row is still written, action = the negative id, done = 0).
     * terminal_*: Env.run appends get_state(p) of EVERY seat after the last step of an episode (env.py:161-164).
     * The fused loop writes those per-seat terminal views into a pool, one row group per finished episode, and
     * terminal_row[t][i] = pool row of the episode that ended in cell [t][i] (-1: no episode ended there).  Rows are
     * handed out with atomicAdd on *terminal_count (in/out, device; zero it to recycle the pool); when the pool is
     * full the row is -1 and err flag 8 is raised. */
    const int32_t *forced_actions;   /* [T][n]                                                               */
    void *terminal_obs;              /* [terminal_capacity][P][obs_stride], obs_dtype                        */
    void *terminal_mask;             /* [terminal_capacity][A] or [terminal_capacity][mask_words]            */
    int32_t *terminal_row;           /* [T][n]                                                               */
    int32_t *terminal_count;         /* [1]                                                                  */
    int32_t terminal_capacity;
} rlc_trajectory;

int rlc_abi_version(void);
const char *rlc_last_error(void);

/* rlcard.make() metadata: Env.num_players/num_actions/state_shape (env.py:41-43) */
int rlc_game_info(int game_id, rlc_info *out);

/* constant rule tables (DouDizhu action table = games/doudizhu/jsondata.zip, utils.py:14-38),
 * uploaded once per device; blob format in DESIGN.md.  RLC_LEDUC / RLC_LIMIT take no blob (NULL, 0): the call tabulates
 * the betting automaton (and for RLC_LIMIT the 7-card evaluator's lookup tables, games/limitholdem/utils.py:37-84,526-614)
 * on the device from the engine itself; without it rlc_rollout_random runs the generic engine (same results). */
int rlc_upload_tables(int game_id, int device, const void *blob, size_t nbytes);

/* Env.reset (env.py:52-63) for envs with reset_mask[i] != 0 (NULL = all): deal, first state,
 * writes obs/mask/cur_player/done=0.  state must be zero-initialised before the first reset. */
int rlc_reset(int game_id, const rlc_buffers *b, int n, const uint8_t *reset_mask, void *stream);

/* Env.step (env.py:65-86) for every env: actions[i] < 0 = leave env i untouched.  Writes the next
 * state's obs/mask/cur_player, done and payoffs (Env.get_payoffs, env.py:199-207). */
int rlc_step(int game_id, const rlc_buffers *b, const int32_t *actions, int n, int flags, void *stream);

/* Env.get_state(player_id) (env.py:188-197): obs/mask as seen by seat[i] (NULL = current player);
 * also refreshes cur_player/done/payoffs. */
int rlc_observe(int game_id, const rlc_buffers *b, const int32_t *seat, int n, void *stream);

/* The Env.run loop with RandomAgents (env.py:120-169, agents/random_agent.py:17-27) fused on
 * device: k_steps env-steps per env, uniform-random legal actions from the Philox policy stream,
 * auto reset.  Trajectory rows as documented above.  Results depend only on (seed, global env id, step
 * index), never on the launch geometry: the thread-per-env games run 32 / 16 / 8 envs per warp depending
 * on the game and n (environment variable RLC_ROLLOUT_EPW forces one, for tuning and tests). */
int rlc_rollout_random(int game_id, const rlc_buffers *b, const rlc_trajectory *traj, int n, int k_steps,
                       void *stream);

/* ---- the judgers / encoders as standalone operators (device pointers; used by the known-answer tests and by
 * callers that score hands outside an episode) ---- */

/* games/limitholdem/utils.py:526-569 compare_hands: cards uint8 [n][P][7], card id = 13*suit + rank as in
 * games/limitholdem/card2index.json (S,H,D,C x A,2..K); first card 255 = folded (None).
 * winners uint8 [n][P]: 1 for every unfolded hand of maximal strength. */
int rlc_judge_holdem(const uint8_t *cards, int n, int num_players, uint8_t *winners, void *stream);

/* games/leducholdem/judger.py:12-64 judge_game / big blind (game.py:170-178): cases int32 [n][7] =
 * (rank0, rank1, public rank or -1, chips0, chips1, folded0, folded1), ranks J,Q,K = 0,1,2 -> payoffs float [n][2] */
int rlc_judge_leduc(const int32_t *cases, int n, float *payoffs, void *stream);

/* games/doudizhu/judger.py:124-258 playable_cards_from_hand (targets NULL or targets[i] < 0: lead) and
 * games/doudizhu/utils.py:225-262 get_gt_cards (targets[i] = action id to beat): hands uint8 [n][15] rank counts
 * (3456789TJQKA2BR) -> bit-packed legal sets uint32 [n][860] (859 words of ids + one pad word).  Needs rlc_upload_tables(RLC_DOUDIZHU, ...). */
int rlc_judge_doudizhu(const uint8_t *hands, const int32_t *targets, int n, uint32_t *mask, void *stream);

/* games/uno/utils.py:86-127 encode_hand + encode_target: hands uint8 [n][32] card codes 15*colour + trait
 * (255 = empty slot), targets uint8 [n] -> obs uint8 [n][240] ([4][4][15] planes of envs/uno.py:24-33) */
int rlc_encode_uno(const uint8_t *hands, const uint8_t *targets, int n, uint8_t *obs, void *stream);

/* ---- the DMC actor's bookkeeping (rlcard/agents/dmc_agent/utils.py:97-155 `act`) ----
 * A window of a rollout trajectory ([T][n] rows, auto reset) is folded into per-position row pools: one row per
 * decision of position p with the state p saw, the feature of the action taken (env.get_action_feature: 54-d for
 * DouDizhu, one-hot over num_actions otherwise), target = p's payoff at the end of that episode, done = p's last
 * decision of the episode, episode_return = payoff when done else 0.  Decisions of episodes still running at
 * the end of the window wait in the per-env open-episode store and are emitted by a later call.  All memory is
 * caller-owned; open_len / out_count / overflow must be zero-initialised before the first call. */
typedef struct rlc_dmc_buffers {
    void *open_obs;                 /* [n][open_capacity][obs row bytes]                                     */
    int32_t *open_action;           /* [n][open_capacity]                                                    */
    int8_t *open_player;            /* [n][open_capacity]                                                    */
    int32_t *open_len;              /* [n]                                                                   */
    int32_t open_capacity;          /* >= longest episode (decisions of all seats)                            */
    void *out_state[RLC_MAX_PLAYERS];            /* [out_capacity][obs row bytes], dtype of the trajectory    */
    int8_t *out_action[RLC_MAX_PLAYERS];         /* [out_capacity][F], F = 54 (doudizhu) or num_actions       */
    float *out_target[RLC_MAX_PLAYERS];          /* [out_capacity]                                            */
    float *out_episode_return[RLC_MAX_PLAYERS];  /* [out_capacity]                                            */
    uint8_t *out_done[RLC_MAX_PLAYERS];          /* [out_capacity]                                            */
    int32_t *out_count;             /* [RLC_MAX_PLAYERS] rows used per position (in/out, device)              */
    int32_t out_capacity;
    int32_t *overflow;              /* [1] device flag: 1 a pool was full (rows dropped), 2 open store too small */
} rlc_dmc_buffers;

int rlc_dmc_collect(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_dmc_buffers *b,
                    void *stream);

/* The DMC agent's view of a state (rlcard/agents/dmc_agent/model.py:91-110 predict): legal ids in ascending order
 * (ids int32 [n][max_ids], -1 padded; count int32 [n], may exceed max_ids if the list was cut) from mask rows in the
 * layout rlc_step writes, and one feature row per action id (env.get_action_feature: 54-d thermometer for doudizhu,
 * one-hot over num_actions otherwise; id < 0 -> zeros), out int8 [m][F]. */
int rlc_legal_ids(int game_id, const void *mask, int n, int max_ids, int32_t *ids, int32_t *count, void *stream);
int rlc_action_features(int game_id, const int32_t *ids, int m, int8_t *out, void *stream);

/* ---- run_rl.py's data path: Env.run(is_training=True) + reorganize (rlcard/utils/utils.py:153-179) as a per-step
 * stream of [state, action, reward, next_state (+ its legal mask), done] rows per seat.
 * Call with phase 0 BEFORE rlc_step (b->obs / mask / cur_player hold the state the acting seat sees, `actions` the
 * ids about to be applied) and with phase 1 AFTER it (done, payoffs, terminal_obs, and the mask of the post-step
 * state; use rlc_step without RLC_AUTO_RESET but with RLC_TERMINAL_OBS, then rlc_reset the finished envs).
 * pend_valid / out_count / overflow must be zero-initialised before the first call. */
typedef struct rlc_rl_buffers {
    void *pend_obs;                 /* [n][P][obs row bytes] last decision of each seat awaiting its next state */
    int32_t *pend_action;           /* [n][P]                                                                 */
    uint8_t *pend_valid;            /* [n][P]                                                                 */
    void *out_state[RLC_MAX_PLAYERS];        /* [out_capacity][obs row bytes]                                  */
    int32_t *out_action[RLC_MAX_PLAYERS];    /* [out_capacity]                                                 */
    float *out_reward[RLC_MAX_PLAYERS];      /* [out_capacity] the seat's payoff on its last transition, else 0 */
    void *out_next_state[RLC_MAX_PLAYERS];   /* [out_capacity][obs row bytes]                                  */
    void *out_next_mask[RLC_MAX_PLAYERS];    /* [out_capacity][mask row bytes] legal set of next_state          */
    uint8_t *out_done[RLC_MAX_PLAYERS];      /* [out_capacity]                                                 */
    int32_t *out_count;             /* [RLC_MAX_PLAYERS] (in/out, device)                                      */
    int32_t out_capacity;
    int32_t *overflow;              /* [1]                                                                     */
} rlc_rl_buffers;

int rlc_rl_feed(int game_id, int phase, const rlc_buffers *env, const int32_t *actions, int n, const rlc_rl_buffers *b,
                void *stream);

/* reorganize (rlcard/utils/utils.py:153-179) over a FUSED rollout window: traj = the [T][n] rows written by
 * rlc_rollout_random with the terminal_* pool enabled (obs, mask, action, player, done, payoffs, terminal_obs,
 * terminal_mask, terminal_row are all required).  Emits the same per-seat transition rows as rlc_rl_feed into the same
 * rlc_rl_buffers; decisions whose next state lies in a later window wait in pend_* (so consecutive windows of one
 * VecEnv chain exactly).  Cells with a negative action id (idle envs of a forced-action replay) are skipped. */
int rlc_reorganize(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, const rlc_rl_buffers *b,
                   void *stream);

/* np.random.RandomState seeding on device (rlcard/utils/seeding.py:33-41 -> RandomState.seed(list) = MT19937
 * init_by_array): key_words uint32 [n][2] (the 1-2 words of sha512(str(seed))[:8], computed by the caller),
 * key_len int32 [n] -> mt uint32 [625][n] in the layout of rlc_buffers.mt (624 state words + index). */
int rlc_seed_mt19937(const uint32_t *key_words, const int32_t *key_len, int n, uint32_t *mt, void *stream);

/* Compact host wire format of a dense trajectory window (csrc/tu_compact.cu has the record layouts): [T][n] cells of
 * obs + mask + action + player + done + payoffs -> out uint32 [T][n][words], words = rlc_compact_words(game): Leduc 1,
 * Limit Hold'em 3, UNO 7, DouDizhu 33, Scout 20 (float32 obs only); 0 = the game has no compact format and
 * rlc_compact_trajectory returns RLC_ENOTIMPL.  A host consumer behind PCIe fetches 4 / 12 / 28 / 132 / 80 bytes per env-step
 * instead of 57 / 93 / 319 / 4374 / 2983 and expands rows on demand (rlcard_b200/compact.py; DouDizhu's legal mask is not sent
 * but recomputed on the host from the row and the action table). */
int rlc_compact_words(int game_id);
int rlc_compact_trajectory(int game_id, const rlc_trajectory *traj, int obs_dtype, int T, int n, uint32_t *out, void *stream);

/* number of kernels this library has launched in this process (bench bookkeeping) */
int64_t rlc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
