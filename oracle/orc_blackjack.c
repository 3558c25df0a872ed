/*
 * oracle/orc_blackjack.c -- CPU ORACLE (test infrastructure, see orc.h).
 * Blackjack, 1 player, 1 deck (envs/blackjack.py:7-10 defaults).
 */
#include "orc.h"
#include <string.h>

typedef struct {
    uint8_t deck[52]; int deck_len;
    uint8_t phand[24], dhand[24]; int np_, nd;
    int winner;      /* game.winner['player0']: 0 ongoing, 2 win, 1 tie, -1 lose (judger.py:25-52) */
    int p_bust, d_bust, p_score, d_score;
} bj_t;

/* games/blackjack/judger.py:54-73 (== envs/blackjack.py:92-103): card id = 13*suit + rank, rank A=0,2..9,T,J,Q,K */
static int bj_score(const uint8_t *h, int n) {
    int score = 0, aces = 0;
    for (int i = 0; i < n; i++) {
        int r = h[i] % 13;
        score += r == 0 ? 11 : (r >= 9 ? 10 : r + 1);
        aces += r == 0;
    }
    while (score > 21 && aces > 0) { aces--; score -= 10; }
    return score;
}
/* games/blackjack/dealer.py:26-37 */
static void bj_deal(bj_t *g, orc_chance *ch, uint8_t *hand, int *n) {
    int idx = (int)orc_below(ch, (uint32_t)g->deck_len);
    hand[(*n)++] = g->deck[idx];
    memmove(g->deck + idx, g->deck + idx + 1, (size_t)(g->deck_len - idx - 1));
    g->deck_len--;
}
static void bj_create(void *s) { (void)s; }
/* games/blackjack/game.py:22-54, dealer.py:6-24 */
static int bj_reset(void *s, orc_chance *ch) {
    bj_t *g = (bj_t *)s;
    for (int i = 0; i < 52; i++) g->deck[i] = (uint8_t)i;
    /* throughput (Philox) spec: deal_card() already draws a uniform index into the remaining deck, so the
     * shuffle adds nothing to the distribution and is not drawn; the deck stays in id order */
    if (ch->kind != ORC_CHANCE_PHILOX) orc_shuffle_u8(ch, g->deck, 52);
    g->deck_len = 52;
    g->np_ = g->nd = 0;
    for (int i = 0; i < 2; i++) { bj_deal(g, ch, g->phand, &g->np_); bj_deal(g, ch, g->dhand, &g->nd); }
    g->p_score = bj_score(g->phand, g->np_); g->p_bust = g->p_score > 21;
    g->d_score = bj_score(g->dhand, g->nd); g->d_bust = g->d_score > 21;
    g->winner = 0;
    return 0;
}
/* dealer plays out + judge (game.py:80-88, 94-102; judger.py:25-52) */
static void bj_finish(bj_t *g, orc_chance *ch) {
    while (bj_score(g->dhand, g->nd) < 17) bj_deal(g, ch, g->dhand, &g->nd);
    g->d_score = bj_score(g->dhand, g->nd); g->d_bust = g->d_score > 21;
    if (g->p_bust) g->winner = -1;
    else if (g->d_bust) g->winner = 2;
    else g->winner = g->p_score > g->d_score ? 2 : (g->p_score < g->d_score ? -1 : 1);
}
/* games/blackjack/game.py:56-123; ids hit 0, stand 1 (envs/blackjack.py:24,81-90) */
static int bj_step(void *s, orc_chance *ch, int id) {
    bj_t *g = (bj_t *)s;
    if (id != 1) {
        bj_deal(g, ch, g->phand, &g->np_);
        g->p_score = bj_score(g->phand, g->np_); g->p_bust = g->p_score > 21;
        if (g->p_bust) bj_finish(g, ch);
    } else {
        g->p_score = bj_score(g->phand, g->np_); g->p_bust = g->p_score > 21;
        bj_finish(g, ch);
    }
    return 0;
}
static int bj_legal(const void *s, uint8_t *mask) { (void)s; mask[0] = mask[1] = 1; return 2; }   /* envs/blackjack.py:55 */
/* envs/blackjack.py:38-60; games/blackjack/game.py:162-190 (dealer.hand[1:] until over) */
static int bj_obs(const void *s, int seat, float *o) {
    const bj_t *g = (const bj_t *)s; (void)seat;
    o[0] = (float)bj_score(g->phand, g->np_);
    o[1] = (float)(g->winner != 0 ? bj_score(g->dhand, g->nd) : bj_score(g->dhand + 1, g->nd - 1));
    return 2;
}
static int bj_over(const void *s) { return ((const bj_t *)s)->winner != 0; }   /* game.py:192-205 */
static int bj_player(const void *s) { (void)s; return 0; }
/* envs/blackjack.py:62-78 */
static void bj_payoffs(const void *s, double *out) {
    int w = ((const bj_t *)s)->winner;
    out[0] = w == 2 ? 1.0 : (w == 1 ? 0.0 : -1.0);
}
const orc_game_vt orc_vt_blackjack = { "blackjack", 1, 2, {2, 0, 0, 0}, sizeof(bj_t), bj_create, bj_reset, bj_step,
    bj_legal, bj_obs, bj_over, bj_player, bj_payoffs };
