/*
 * oracle/orc_nolimit.c -- CPU ORACLE (test infrastructure, see orc.h).
 * No-limit Texas Hold'em, 2 players, 100 chips each, 5 actions (games/nolimitholdem/{game,round,player}.py on
 * top of games/limitholdem/{dealer,player,judger,utils}.py; envs/nolimitholdem.py).
 */
#include "orc.h"
#include <string.h>

enum { N_FOLD = 0, N_CHECK_CALL = 1, N_HALF_POT = 2, N_POT = 3, N_ALL_IN = 4 };   /* round.py:8-18 Action */
enum { ST_ALIVE = 0, ST_FOLDED = 1, ST_ALLIN = 2 };                               /* limitholdem/player.py:4-7 */

typedef struct {
    uint8_t deck[52]; int deck_len;
    int hand[2][2], n_public, public_cards[5];
    int in_chips[2], remained[2], status[2];
    int dealer_id;                       /* -1 until the first init_game draws it; then kept (game.py:62-63) */
    int game_pointer, round_counter, pot;
    int raised[2], not_raise_num, not_playing_num;      /* round.py:34-46 */
    int created;
} nolimit_t;

static void nl_create(void *s) { nolimit_t *g = (nolimit_t *)s; memset(g, 0, sizeof *g); g->dealer_id = -1; g->created = 1; }
static int nl_deal(nolimit_t *g) { return g->deck[--g->deck_len]; }
static void nl_bet(nolimit_t *g, int p, int chips) {                  /* nolimitholdem/player.py:16-19 */
    int q = chips <= g->remained[p] ? chips : g->remained[p];
    g->in_chips[p] += q; g->remained[p] -= q;
}
static int imax(int a, int b) { return a > b ? a : b; }

/* games/nolimitholdem/game.py:50-111 */
static int nl_reset(void *s, orc_chance *ch) {
    nolimit_t *g = (nolimit_t *)s;
    if (!g->created) { g->dealer_id = -1; g->created = 1; }
    if (g->dealer_id < 0) g->dealer_id = (int)orc_below(ch, 2);         /* np_random.randint(0, num_players) */
    for (int i = 0; i < 52; i++) g->deck[i] = (uint8_t)i;
    if (ch->kind == ORC_CHANCE_PHILOX) {                 /* throughput spec: three words deal the nine cards ever popped */
        uint32_t x1 = orc_below(ch, 52u * 51u * 50u), x2 = orc_below(ch, 49u * 48u * 47u), x3 = orc_below(ch, 46u * 45u * 44u);
        int j[9] = { (int)(x1 / 2550), (int)((x1 / 50) % 51), (int)(x1 % 50), (int)(x2 / 2256), (int)((x2 / 47) % 48),
                     (int)(x2 % 47), (int)(x3 / 1980), (int)((x3 / 44) % 45), (int)(x3 % 44) };
        for (int k = 0; k < 9; k++) { uint8_t t = g->deck[51 - k]; g->deck[51 - k] = g->deck[j[k]]; g->deck[j[k]] = t; }
    } else orc_shuffle_tail_u8(ch, g->deck, 52, 9);
    g->deck_len = 52;
    for (int p = 0; p < 2; p++) { g->in_chips[p] = 0; g->remained[p] = 100; g->status[p] = ST_ALIVE; }
    for (int i = 0; i < 4; i++) g->hand[i % 2][i / 2] = nl_deal(g);
    g->n_public = 0;
    int sb = (g->dealer_id + 1) % 2, bb = (g->dealer_id + 2) % 2;
    nl_bet(g, bb, 2); nl_bet(g, sb, 1);
    g->game_pointer = (bb + 1) % 2;
    g->not_raise_num = 0; g->not_playing_num = 0;
    g->raised[0] = g->in_chips[0]; g->raised[1] = g->in_chips[1];
    g->round_counter = 0;
    g->pot = g->in_chips[0] + g->in_chips[1];                           /* get_state, game.py:186 */
    return g->game_pointer;
}
/* games/nolimitholdem/round.py:125-161 -> bit a set when action a is legal */
static int nl_legal_bits(const nolimit_t *g) {
    int p = g->game_pointer, m = 0x1F, mx = imax(g->raised[0], g->raised[1]);
    int diff = mx - g->raised[p];
    if (diff > 0 && diff >= g->remained[p]) m &= ~((1 << N_HALF_POT) | (1 << N_POT) | (1 << N_ALL_IN));
    else {
        if (g->pot > g->remained[p]) m &= ~(1 << N_POT);
        if (g->pot / 2 > g->remained[p]) m &= ~(1 << N_HALF_POT);
        if ((m >> N_HALF_POT & 1) && g->pot / 2 + g->raised[p] <= mx) m &= ~(1 << N_HALF_POT);
    }
    return m;
}
/* env.py:65-86, envs/nolimitholdem.py:91-105, game.py:113-181, round.py:65-123 */
static int nl_step(void *s, orc_chance *ch, int id) {
    nolimit_t *g = (nolimit_t *)s; (void)ch;
    int legal = nl_legal_bits(g);
    if (id < 0 || id > 4 || !((legal >> id) & 1)) id = N_FOLD;          /* the reference crashes here (Action.CHECK does not exist) */
    int p = g->game_pointer, mx = imax(g->raised[0], g->raised[1]);
    switch (id) {                                                       /* round.py:77-101 */
    case N_CHECK_CALL: { int diff = mx - g->raised[p]; g->raised[p] = mx; nl_bet(g, p, diff); g->not_raise_num++; break; }
    case N_ALL_IN: { int q = g->remained[p]; g->raised[p] += q; nl_bet(g, p, q); g->not_raise_num = 1; break; }
    case N_POT: g->raised[p] += g->pot; nl_bet(g, p, g->pot); g->not_raise_num = 1; break;
    case N_HALF_POT: { int q = g->pot / 2; g->raised[p] += q; nl_bet(g, p, q); g->not_raise_num = 1; break; }
    case N_FOLD: g->status[p] = ST_FOLDED; break;
    }
    if (g->remained[p] == 0 && g->status[p] != ST_FOLDED) g->status[p] = ST_ALLIN;
    g->game_pointer = (p + 1) % 2;
    if (g->status[p] == ST_ALLIN) { g->not_playing_num++; g->not_raise_num--; }
    if (g->status[p] == ST_FOLDED) g->not_playing_num++;
    for (int k = 0; k < 2 && g->status[g->game_pointer] == ST_FOLDED; k++) g->game_pointer = (g->game_pointer + 1) % 2;
    /* game.py:141-147 */
    int bypass[2] = { g->status[0] != ST_ALIVE, g->status[1] != ST_ALIVE };
    if (2 - (bypass[0] + bypass[1]) == 1) {
        int last = bypass[0] ? 1 : 0;
        if (g->raised[last] >= imax(g->raised[0], g->raised[1])) bypass[last] = 1;
    }
    if (g->not_raise_num + g->not_playing_num >= 2) {                   /* round.is_over, game.py:150-177 */
        int all = bypass[0] + bypass[1] == 2;
        g->game_pointer = (g->dealer_id + 1) % 2;
        if (!all) while (bypass[g->game_pointer]) g->game_pointer = (g->game_pointer + 1) % 2;
        if (g->round_counter == 0) { for (int k = 0; k < 3; k++) g->public_cards[g->n_public++] = nl_deal(g); if (all) g->round_counter++; }
        if (g->round_counter == 1) { g->public_cards[g->n_public++] = nl_deal(g); if (all) g->round_counter++; }
        if (g->round_counter == 2) { g->public_cards[g->n_public++] = nl_deal(g); if (all) g->round_counter++; }
        g->round_counter++;
        g->not_raise_num = 0; g->raised[0] = g->raised[1] = 0;          /* start_new_round */
    }
    g->pot = g->in_chips[0] + g->in_chips[1];                           /* get_state */
    return g->game_pointer;
}
static int nl_legal(const void *s, uint8_t *mask) {
    int m = nl_legal_bits((const nolimit_t *)s), c = 0;
    for (int a = 0; a < 5; a++) { mask[a] = (uint8_t)((m >> a) & 1); c += mask[a]; }
    return c;
}
/* envs/nolimitholdem.py:47-79 */
static int nl_obs(const void *s, int seat, float *o) {
    const nolimit_t *g = (const nolimit_t *)s;
    if (seat < 0) seat = g->game_pointer;
    memset(o, 0, 54 * sizeof(float));
    for (int i = 0; i < g->n_public; i++) o[g->public_cards[i]] = 1.f;
    o[g->hand[seat][0]] = 1.f; o[g->hand[seat][1]] = 1.f;
    o[52] = (float)g->in_chips[seat];
    o[53] = (float)imax(g->in_chips[0], g->in_chips[1]);
    return 54;
}
/* limitholdem/game.py:216-231 (inherited) */
static int nl_over(const void *s) {
    const nolimit_t *g = (const nolimit_t *)s;
    return ((g->status[0] != ST_FOLDED) + (g->status[1] != ST_FOLDED) == 1) || g->round_counter >= 4;
}
static int nl_player(const void *s) { return ((const nolimit_t *)s)->game_pointer; }
/* game.py:226-236 + limitholdem/judger.py:11-108 for two players: the winner takes what the loser can match */
static void nl_payoffs(const void *s, double *out) {
    const nolimit_t *g = (const nolimit_t *)s;
    int has[2] = { g->status[0] != ST_FOLDED, g->status[1] != ST_FOLDED }, w[2];
    if (has[0] != has[1]) { w[0] = has[0]; w[1] = has[1]; }
    else {
        uint32_t str[2];
        for (int p = 0; p < 2; p++) {
            uint8_t c[7] = { (uint8_t)g->hand[p][0], (uint8_t)g->hand[p][1] };
            for (int k = 0; k < 5; k++) c[2 + k] = (uint8_t)g->public_cards[k];
            str[p] = orc_holdem_strength7(c);
        }
        w[0] = str[0] >= str[1]; w[1] = str[1] >= str[0];
    }
    /* split_pots_among_players: first pot = min(in_chips) from each, the rest goes back to its owner */
    int m = g->in_chips[0] < g->in_chips[1] ? g->in_chips[0] : g->in_chips[1];
    out[0] = out[1] = 0;
    if (w[0] != w[1]) { out[0] = w[0] ? m : -m; out[1] = -out[0]; }
}
const orc_game_vt orc_vt_nolimit = { "no-limit-holdem", 2, 5, {54, 54, 0, 0}, sizeof(nolimit_t), nl_create, nl_reset, nl_step,
    nl_legal, nl_obs, nl_over, nl_player, nl_payoffs };
