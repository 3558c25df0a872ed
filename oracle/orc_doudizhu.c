/*
 * oracle/orc_doudizhu.c -- CPU ORACLE (test infrastructure, see orc.h).
 * DouDizhu, 3 players, 27 472 actions (games/doudizhu/{game,round,player,dealer,judger,utils}.py,
 * envs/doudizhu.py).  Hands are rank counts (suits never matter).  Legal sets follow the reference
 * semantics restated over the action table: lead = every non-pass action contained in the hand
 * (== judger.playable_cards_from_hand, SURVEY F-DDZ1), follow = utils.get_gt_cards.
 * The action table (id -> rank counts, type, weight) is installed by orc_doudizhu_set_table().
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>

#define DDZ_A 27472
#define DDZ_PASS 27471
static uint8_t (*T_counts)[15];
static uint8_t *T_type, *T_weight, *T_len;
static int T_bomb = -1, T_rocket = -1;

void orc_doudizhu_set_table(const uint64_t *packed, const uint8_t *type, const uint8_t *weight, int n, int bomb_type, int rocket_type) {
    if (n != DDZ_A) return;
    if (!T_counts) { T_counts = malloc(sizeof(uint8_t[15]) * DDZ_A); T_type = malloc(DDZ_A); T_weight = malloc(DDZ_A); T_len = malloc(DDZ_A); }
    for (int i = 0; i < n; i++) {
        int len = 0;
        for (int r = 0; r < 15; r++) { T_counts[i][r] = (uint8_t)((packed[i] >> (4 * r)) & 15); len += T_counts[i][r]; }
        T_type[i] = type[i]; T_weight[i] = weight[i]; T_len[i] = (uint8_t)len;
    }
    T_bomb = bomb_type; T_rocket = rocket_type;
}

typedef struct {
    int hand[3][15], played[3][15];
    int trace_p[512], trace_a[512], ntrace;
    int greater, greater_action, cur, winner;
} ddz_t;

static int ddz_contains(const int *hand, int a) {                    /* utils.py:158-193 contains_cards */
    for (int r = 0; r < 15; r++) if (T_counts[a][r] > hand[r]) return 0;
    return 1;
}
static int hand_size(const int *h) { int n = 0; for (int r = 0; r < 15; r++) n += h[r]; return n; }

/* player.py:60-76 + game.py:110-128: the current player's actions (empty once the game is over) */
static int ddz_legal_list(const ddz_t *g, uint8_t *mask) {
    int cnt = 0;
    if (mask) memset(mask, 0, DDZ_A);
    if (g->winner >= 0) return 0;
    const int *hand = g->hand[g->cur];
    int n = hand_size(hand);
    if (g->greater < 0 || g->greater == g->cur) {                  /* lead: judger.get_playable_cards */
        for (int a = 0; a < DDZ_PASS; a++)
            if (T_len[a] <= n && ddz_contains(hand, a)) { if (mask) mask[a] = 1; cnt++; }
    } else {                                                        /* follow: utils.py:225-262 get_gt_cards */
        int tt = T_type[g->greater_action], tw = T_weight[g->greater_action];
        if (mask) mask[DDZ_PASS] = 1;
        cnt++;
        if (tt == T_rocket) return cnt;
        for (int a = 0; a < DDZ_PASS; a++) {
            int t = T_type[a], ok;
            if (t == tt) ok = T_weight[a] > tw;
            else ok = (t == T_bomb) || (t == T_rocket);             /* type_dict['bomb'] = -1 unless the target is a bomb */
            if (ok && T_len[a] <= n && ddz_contains(hand, a)) { if (mask) mask[a] = 1; cnt++; }
        }
    }
    return cnt;
}
static void ddz_create(void *s) { (void)s; }
/* game.py:23-51, round.py:25-39, dealer.py:12-76: rank-sorted 54 deck, one shuffle, 17/17/17 + 3 to seat 0 */
static int ddz_reset(void *s, orc_chance *ch) {
    ddz_t *g = (ddz_t *)s;
    uint8_t deck[54];
    memset(g, 0, sizeof *g);
    for (int i = 0; i < 52; i++) deck[i] = (uint8_t)(i / 4);
    deck[52] = 13; deck[53] = 14;
    orc_shuffle_u8(ch, deck, 54);
    for (int p = 0; p < 3; p++) for (int k = 0; k < 17; k++) g->hand[p][deck[17 * p + k]]++;
    for (int k = 51; k < 54; k++) g->hand[0][deck[k]]++;
    g->greater = -1; g->greater_action = DDZ_PASS; g->cur = 0; g->winner = -1;
    return 0;
}
/* env.py:65-86, game.py:53-81, round.py:52-79, player.py:78-108, judger.py:335-348 */
static int ddz_step(void *s, orc_chance *ch, int id) {
    ddz_t *g = (ddz_t *)s; (void)ch;
    if (g->winner >= 0) return g->cur;
    if (id < 0 || id >= DDZ_A) id = DDZ_PASS;
    int p = g->cur;
    if (g->ntrace < 512) { g->trace_p[g->ntrace] = p; g->trace_a[g->ntrace] = id; }
    g->ntrace++;
    if (id != DDZ_PASS) {
        for (int r = 0; r < 15; r++) { g->played[p][r] += T_counts[id][r]; g->hand[p][r] -= T_counts[id][r]; if (g->hand[p][r] < 0) g->hand[p][r] = 0; }
        g->greater = p; g->greater_action = id;
    }
    if (hand_size(g->hand[p]) == 0) g->winner = p;
    g->cur = (p + 1) % 3;
    return g->cur;
}
static int ddz_legal(const void *s, uint8_t *mask) { return ddz_legal_list((const ddz_t *)s, mask); }
/* envs/doudizhu.py:153-167 _cards2array on rank counts */
static void enc54(float *o, const int *c) {
    for (int r = 0; r < 13; r++) for (int k = 0; k < 4; k++) o[4 * r + k] = k < c[r] ? 1.f : 0.f;
    o[52] = c[13] ? 1.f : 0.f; o[53] = c[14] ? 1.f : 0.f;
}
static void enc54_action(float *o, int a) {
    int c[15];
    for (int r = 0; r < 15; r++) c[r] = (a >= 0 && a < DDZ_A) ? T_counts[a][r] : 0;
    enc54(o, c);
}
static void one_hot(float *o, int n, int size) {                  /* :169-173, index n-1 (wraps at 0, Q-DDZ1) */
    memset(o, 0, sizeof(float) * (size_t)size);
    o[n >= 1 ? n - 1 : size - 1] = 1.f;
}
/* envs/doudizhu.py:26-91 */
static int ddz_obs(const void *s, int seat, float *o) {
    const ddz_t *g = (const ddz_t *)s;
    if (seat < 0) seat = g->cur;
    int nt = g->ntrace < 512 ? g->ntrace : 512, others[15], left[3];
    for (int r = 0; r < 15; r++) others[r] = g->hand[(seat + 1) % 3][r] + g->hand[(seat + 2) % 3][r];
    for (int p = 0; p < 3; p++) left[p] = hand_size(g->hand[p]);
    float *q = o;
    enc54(q, g->hand[seat]); q += 54;
    enc54(q, others); q += 54;
    int last = -1;
    if (nt > 0) { last = g->trace_a[nt - 1]; if (last == DDZ_PASS) last = nt >= 2 ? g->trace_a[nt - 2] : -1; }
    enc54_action(q, last); q += 54;
    for (int k = 0; k < 9; k++) { int idx = nt - 9 + k; enc54_action(q, idx >= 0 ? g->trace_a[idx] : -1); q += 54; }
    if (seat == 0) {
        enc54(q, g->played[2]); q += 54;
        enc54(q, g->played[1]); q += 54;
        one_hot(q, left[2], 17); q += 17;
        one_hot(q, left[1], 17); q += 17;
        return 790;
    }
    int mate = 3 - seat, last_ll = -1, last_mate = -1;
    for (int i = nt - 1; i >= 0; i--) if (g->trace_p[i] == 0) { last_ll = g->trace_a[i]; break; }   /* reference raises if none */
    for (int i = nt - 1; i >= 0; i--) if (g->trace_p[i] == mate) { last_mate = g->trace_a[i]; break; }
    enc54(q, g->played[0]); q += 54;
    enc54(q, g->played[mate]); q += 54;
    enc54_action(q, last_ll); q += 54;
    enc54_action(q, last_mate); q += 54;
    one_hot(q, left[0], 20); q += 20;
    one_hot(q, left[mate], 17); q += 17;
    return 901;
}
static int ddz_over(const void *s) { return ((const ddz_t *)s)->winner >= 0; }
static int ddz_player(const void *s) { return ((const ddz_t *)s)->cur; }
static void ddz_payoffs(const void *s, double *out) {              /* judger.py:351-359, landlord = seat 0 */
    const ddz_t *g = (const ddz_t *)s;
    out[0] = g->winner == 0 ? 1 : 0; out[1] = out[2] = g->winner == 0 ? 0 : 1;
}
/* known-answer helper: legal set of an arbitrary hand (rank counts), leading (target < 0 ==
 * judger.playable_cards_from_hand) or following `target` (== utils.get_gt_cards); returns the count */
int orc_doudizhu_legal_for(const uint8_t *hand15, int target_action, uint8_t *mask) {
    ddz_t g;
    memset(&g, 0, sizeof g);
    for (int r = 0; r < 15; r++) g.hand[0][r] = hand15[r];
    g.cur = 0; g.winner = -1;
    if (target_action < 0) { g.greater = -1; g.greater_action = DDZ_PASS; }
    else { g.greater = 1; g.greater_action = target_action; }
    return ddz_legal_list(&g, mask);
}
const orc_game_vt orc_vt_doudizhu = { "doudizhu", 3, DDZ_A, {790, 901, 901, 0}, sizeof(ddz_t), ddz_create, ddz_reset, ddz_step,
    ddz_legal, ddz_obs, ddz_over, ddz_player, ddz_payoffs };
