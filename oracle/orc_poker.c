/*
 * oracle/orc_poker.c -- CPU ORACLE (test infrastructure, see orc.h).
 * Leduc Hold'em and Limit Texas Hold'em (2 players), sharing the betting round
 * of games/limitholdem/round.py, plus the 7-card evaluator.
 */
#include "orc.h"
#include <string.h>

enum { A_CALL = 0, A_RAISE = 1, A_FOLD = 2, A_CHECK = 3 };   /* envs/leducholdem.py:26, envs/limitholdem.py:25 */

/* ---- games/limitholdem/round.py (shared by both games) ---- */
typedef struct {
    int game_pointer, raise_amount, allowed_raise_num, have_raised, not_raise_num, raised[2];
} round_t;

static int imax2(const int *r) { return r[0] > r[1] ? r[0] : r[1]; }

/* round.py:35-51 */
static void round_start(round_t *r, int pointer, const int *raised) {
    r->game_pointer = pointer; r->have_raised = 0; r->not_raise_num = 0;
    r->raised[0] = raised ? raised[0] : 0; r->raised[1] = raised ? raised[1] : 0;
}
/* round.py:95-116 -> bit a set when action id a is legal */
static int round_legal(const round_t *r) {
    int m = 0xF, mx = imax2(r->raised);
    if (r->have_raised >= r->allowed_raise_num) m &= ~(1 << A_RAISE);
    if (r->raised[r->game_pointer] < mx) m &= ~(1 << A_CHECK);
    if (r->raised[r->game_pointer] == mx) m &= ~(1 << A_CALL);
    return m;
}
/* round.py:53-93 */
static int round_proceed(round_t *r, int *in_chips, int *folded, int action) {
    int p = r->game_pointer, mx = imax2(r->raised), diff;
    switch (action) {
    case A_CALL: diff = mx - r->raised[p]; r->raised[p] = mx; in_chips[p] += diff; r->not_raise_num++; break;
    case A_RAISE: diff = mx - r->raised[p] + r->raise_amount; r->raised[p] = mx + r->raise_amount;
        in_chips[p] += diff; r->have_raised++; r->not_raise_num = 1; break;
    case A_FOLD: folded[p] = 1; break;
    case A_CHECK: r->not_raise_num++; break;
    }
    r->game_pointer = (p + 1) % 2;
    for (int k = 0; k < 2 && folded[r->game_pointer]; k++) r->game_pointer = (r->game_pointer + 1) % 2;
    return r->game_pointer;
}
static int round_over(const round_t *r) { return r->not_raise_num >= 2; }   /* round.py:118-127 */
/* envs/leducholdem.py:81-96, envs/limitholdem.py:81-96: illegal id -> check, else fold */
static int decode_action(const round_t *r, int id) {
    int m = round_legal(r);
    if (id < 0 || id > 3 || !((m >> id) & 1)) return ((m >> A_CHECK) & 1) ? A_CHECK : A_FOLD;
    return id;
}
static int legal_to_mask(int m, uint8_t *mask) {
    int c = 0;
    for (int a = 0; a < 4; a++) { mask[a] = (uint8_t)((m >> a) & 1); c += mask[a]; }
    return c;
}

/* =============================================================== Leduc */
typedef struct {
    uint8_t deck[6]; int deck_len;
    int hand[2], in_chips[2], folded[2], public_card, game_pointer, round_counter;
    round_t round;
} leduc_t;

static void leduc_create(void *s) { (void)s; }
/* games/leducholdem/game.py:46-95, dealer.py:6-12 (deck order SJ HJ SQ HQ SK HK -> rank = index/2) */
static int leduc_reset(void *s, orc_chance *ch) {
    leduc_t *g = (leduc_t *)s;
    for (int i = 0; i < 6; i++) g->deck[i] = (uint8_t)i;
    if (ch->kind == ORC_CHANCE_PHILOX) {                 /* throughput spec: the three cards ride on the step's base word */
        uint32_t x = orc_chain(ch, 120);
        int j[3] = { (int)(x / 20), (int)((x / 4) % 5), (int)(x % 4) };
        for (int s = 0; s < 3; s++) { uint8_t t = g->deck[5 - s]; g->deck[5 - s] = g->deck[j[s]]; g->deck[j[s]] = t; }
    } else orc_shuffle_tail_u8(ch, g->deck, 6, 3);
    g->deck_len = 6;
    for (int i = 0; i < 2; i++) { g->hand[i] = g->deck[--g->deck_len]; g->in_chips[i] = 0; g->folded[i] = 0; }
    int sb = (int)orc_chain(ch, 2), bb = (sb + 1) % 2;
    g->in_chips[bb] = 2; g->in_chips[sb] = 1;
    g->public_card = -1; g->game_pointer = sb;
    g->round.raise_amount = 2; g->round.allowed_raise_num = 2;
    round_start(&g->round, g->game_pointer, g->in_chips);
    g->round_counter = 0;
    return g->game_pointer;
}
/* envs/env.py:65-86 + games/leducholdem/game.py:97-136 */
static int leduc_step(void *s, orc_chance *ch, int id) {
    leduc_t *g = (leduc_t *)s; (void)ch;
    int action = decode_action(&g->round, id);
    g->game_pointer = round_proceed(&g->round, g->in_chips, g->folded, action);
    if (round_over(&g->round)) {
        if (g->round_counter == 0) { g->public_card = g->deck[--g->deck_len]; g->round.raise_amount = 4; }
        g->round_counter++;
        round_start(&g->round, g->game_pointer, NULL);
    }
    return g->game_pointer;
}
static int leduc_legal(const void *s, uint8_t *mask) { return legal_to_mask(round_legal(&((const leduc_t *)s)->round), mask); }
/* envs/leducholdem.py:41-71 */
static int leduc_obs(const void *s, int seat, float *o) {
    const leduc_t *g = (const leduc_t *)s;
    if (seat < 0) seat = g->game_pointer;
    memset(o, 0, 36 * sizeof(float));
    o[g->hand[seat] >> 1] = 1.f;
    if (g->public_card >= 0) o[3 + (g->public_card >> 1)] = 1.f;
    o[6 + g->in_chips[seat]] = 1.f;
    o[21 + g->in_chips[0] + g->in_chips[1] - g->in_chips[seat]] = 1.f;
    return 36;
}
/* games/leducholdem/game.py:154-168 */
static int leduc_over(const void *s) {
    const leduc_t *g = (const leduc_t *)s;
    return ((!g->folded[0]) + (!g->folded[1]) == 1) || g->round_counter >= 2;
}
static int leduc_player(const void *s) { return ((const leduc_t *)s)->game_pointer; }
/* games/leducholdem/judger.py:12-64 + game.py:170-178 */
static void leduc_payoffs(const void *s, double *out) {
    const leduc_t *g = (const leduc_t *)s;
    int winners[2] = {0, 0}, fold_count = 0, alive_idx = 0, ranks[2];
    for (int i = 0; i < 2; i++) {
        ranks[i] = 11 + (g->hand[i] >> 1);                      /* rank2int J,Q,K */
        if (g->folded[i]) fold_count++; else alive_idx = i;
    }
    if (fold_count == 1) winners[alive_idx] = 1;
    if (winners[0] + winners[1] < 1 && g->public_card >= 0)
        for (int i = 0; i < 2; i++) if ((g->hand[i] >> 1) == (g->public_card >> 1)) { winners[i] = 1; break; }
    if (winners[0] + winners[1] < 1) {
        int mx = ranks[0] > ranks[1] ? ranks[0] : ranks[1];
        for (int i = 0; i < 2; i++) if (ranks[i] == mx) winners[i] = 1;
    }
    double total = g->in_chips[0] + g->in_chips[1], each = total / (winners[0] + winners[1]);
    for (int i = 0; i < 2; i++) out[i] = (winners[i] ? each - g->in_chips[i] : -(double)g->in_chips[i]) / 2.0;
}
/* known-answer helper: LeducholdemJudger.judge_game / big blind on an arbitrary showdown
 * (ranks 0..2 = J,Q,K; pub < 0 = no public card) */
void orc_leduc_judge(int r0, int r1, int pub, int c0, int c1, int f0, int f1, double *out) {
    leduc_t g;
    memset(&g, 0, sizeof g);
    g.hand[0] = 2 * r0; g.hand[1] = 2 * r1 + 1; g.public_card = pub < 0 ? -1 : 2 * pub + (pub == r0 ? 1 : 0);
    g.in_chips[0] = c0; g.in_chips[1] = c1; g.folded[0] = f0; g.folded[1] = f1;
    leduc_payoffs(&g, out);
}
const orc_game_vt orc_vt_leduc = { "leduc-holdem", 2, 4, {36, 36, 0, 0}, sizeof(leduc_t), leduc_create, leduc_reset,
    leduc_step, leduc_legal, leduc_obs, leduc_over, leduc_player, leduc_payoffs };

/* =============================================================== 7-card evaluator */
/* Textbook best-5-of-7 strength; orders hands exactly like games/limitholdem/utils.py
 * (categories :37-84, tie-break key positions :571-614).  Pinned by
 * tests/test_oracle_golden.py against tests/utils/test_holdem_utils.py's KATs and against
 * showdowns recorded from the live compare_hands. */
static int top_straight(unsigned rb) {           /* rb bit r set for rank r (2..14); returns top rank or 0 */
    if (rb & (1u << 14)) rb |= 2u;               /* ace plays low (utils.py:166-182) */
    for (int hi = 14; hi >= 5; hi--) { unsigned need = 0x1Fu << (hi - 4); if ((rb & need) == need) return hi; }
    return 0;
}
uint32_t orc_holdem_strength7(const uint8_t cards[7]) {
    int cnt[15] = {0}, suitcnt[4] = {0}; unsigned rb = 0, srb[4] = {0, 0, 0, 0};
    for (int i = 0; i < 7; i++) {
        int s = cards[i] / 13, ri = cards[i] % 13, r = ri == 0 ? 14 : ri + 1;
        cnt[r]++; suitcnt[s]++; rb |= 1u << r; srb[s] |= 1u << r;
    }
    int fs = -1;
    for (int s = 0; s < 4; s++) if (suitcnt[s] >= 5) fs = s;
    uint32_t kick = 0; int n = 0;
#define PUSH(r) do { kick = (kick << 4) | (uint32_t)(r); n++; } while (0)
#define DONE(cat) do { while (n < 5) { kick <<= 4; n++; } return ((uint32_t)(cat) << 20) | kick; } while (0)
    if (fs >= 0) { int t = top_straight(srb[fs]); if (t) { PUSH(t); DONE(9); } }
    int quad = 0, trips[3] = {0, 0, 0}, nt = 0, pairs[4] = {0, 0, 0, 0}, npair = 0;
    for (int r = 14; r >= 2; r--) {
        if (cnt[r] == 4) quad = r;
        else if (cnt[r] == 3) trips[nt++] = r;
        else if (cnt[r] == 2) pairs[npair++] = r;
    }
    if (quad) { PUSH(quad); for (int r = 14; r >= 2; r--) if (r != quad && cnt[r]) { PUSH(r); break; } DONE(8); }
    if (nt >= 1 && (nt >= 2 || npair >= 1)) {
        int pr = nt >= 2 ? (trips[1] > pairs[0] ? trips[1] : pairs[0]) : pairs[0];
        PUSH(trips[0]); PUSH(pr); DONE(7);
    }
    if (fs >= 0) { for (int r = 14; r >= 2 && n < 5; r--) if (srb[fs] & (1u << r)) PUSH(r); DONE(6); }
    { int t = top_straight(rb); if (t) { PUSH(t); DONE(5); } }
    if (nt == 1) { PUSH(trips[0]); for (int r = 14; r >= 2 && n < 3; r--) if (cnt[r] && r != trips[0]) PUSH(r); DONE(4); }
    if (npair >= 2) {
        PUSH(pairs[0]); PUSH(pairs[1]);
        for (int r = 14; r >= 2; r--) if (cnt[r] && r != pairs[0] && r != pairs[1]) { PUSH(r); break; }
        DONE(3);
    }
    if (npair == 1) { PUSH(pairs[0]); for (int r = 14; r >= 2 && n < 4; r--) if (cnt[r] && r != pairs[0]) PUSH(r); DONE(2); }
    for (int r = 14; r >= 2 && n < 5; r--) if (cnt[r]) PUSH(r);
    DONE(1);
#undef PUSH
#undef DONE
}

/* =============================================================== Limit Hold'em */
typedef struct {
    uint8_t deck[52]; int deck_len;
    int hand[2][2], n_public, public_cards[5], in_chips[2], folded[2], game_pointer, round_counter;
    int raise_nums[4];        /* history_raise_nums of this episode (game.py:101,133) */
    int shown_raise_nums[4];  /* the list the state returned by reset() still points to (Q-LH1, game.py:98 vs :101) */
    int fresh_reset;
    round_t round;
} limit_t;

static void limit_create(void *s) { memset(s, 0, sizeof(limit_t)); }   /* game.py:27 fresh env: zeros */
/* games/limitholdem/game.py:46-103 */
static int limit_reset(void *s, orc_chance *ch) {
    limit_t *g = (limit_t *)s;
    memcpy(g->shown_raise_nums, g->raise_nums, sizeof g->raise_nums);   /* stale list seen by the reset() state */
    for (int i = 0; i < 52; i++) g->deck[i] = (uint8_t)i;               /* utils/utils.py:34-43 == card2index.json */
    int sb;
    if (ch->kind == ORC_CHANCE_PHILOX) {                 /* throughput spec: three deal words (keyed by the episode ordinal) deal nine cards + blind */
        uint32_t x1 = orc_deal_below(ch, 0u, 52u * 51u * 50u), x2 = orc_deal_below(ch, 1u, 49u * 48u * 47u), x3 = orc_deal_below(ch, 2u, 46u * 45u * 44u * 2u);
        uint32_t z = x3 >> 1;
        int j[9] = { (int)(x1 / 2550), (int)((x1 / 50) % 51), (int)(x1 % 50), (int)(x2 / 2256), (int)((x2 / 47) % 48),
                     (int)(x2 % 47), (int)(z / 1980), (int)((z / 44) % 45), (int)(z % 44) };
        for (int s = 0; s < 9; s++) { uint8_t t = g->deck[51 - s]; g->deck[51 - s] = g->deck[j[s]]; g->deck[j[s]] = t; }
        sb = (int)(x3 & 1u);
    } else { orc_shuffle_tail_u8(ch, g->deck, 52, 9); sb = -1; }
    g->deck_len = 52;
    for (int i = 0; i < 4; i++) g->hand[i % 2][i / 2] = g->deck[--g->deck_len];
    g->n_public = 0;
    if (sb < 0) sb = (int)orc_below(ch, 2);
    int bb = (sb + 1) % 2;
    g->in_chips[0] = g->in_chips[1] = 0; g->folded[0] = g->folded[1] = 0;
    g->in_chips[bb] = 2; g->in_chips[sb] = 1;
    g->game_pointer = (bb + 1) % 2;
    g->round.raise_amount = 2; g->round.allowed_raise_num = 4;
    round_start(&g->round, g->game_pointer, g->in_chips);
    g->round_counter = 0;
    memset(g->raise_nums, 0, sizeof g->raise_nums);
    g->fresh_reset = 1;
    return g->game_pointer;
}
/* games/limitholdem/game.py:105-156 */
static int limit_step(void *s, orc_chance *ch, int id) {
    limit_t *g = (limit_t *)s; (void)ch;
    int action = decode_action(&g->round, id);
    g->fresh_reset = 0;
    g->game_pointer = round_proceed(&g->round, g->in_chips, g->folded, action);
    g->raise_nums[g->round_counter] = g->round.have_raised;
    if (round_over(&g->round)) {
        if (g->round_counter == 0) for (int k = 0; k < 3; k++) g->public_cards[g->n_public++] = g->deck[--g->deck_len];
        else if (g->round_counter <= 2) g->public_cards[g->n_public++] = g->deck[--g->deck_len];
        if (g->round_counter == 1) g->round.raise_amount = 4;
        g->round_counter++;
        round_start(&g->round, g->game_pointer, NULL);
    }
    return g->game_pointer;
}
static int limit_legal(const void *s, uint8_t *mask) { return legal_to_mask(round_legal(&((const limit_t *)s)->round), mask); }
/* envs/limitholdem.py:40-71 */
static int limit_obs(const void *s, int seat, float *o) {
    const limit_t *g = (const limit_t *)s;
    const int *rn = g->raise_nums;
    if (seat < 0) { seat = g->game_pointer; if (g->fresh_reset) rn = g->shown_raise_nums; }
    memset(o, 0, 72 * sizeof(float));
    for (int i = 0; i < g->n_public; i++) o[g->public_cards[i]] = 1.f;
    o[g->hand[seat][0]] = 1.f; o[g->hand[seat][1]] = 1.f;
    for (int i = 0; i < 4; i++) o[52 + 5 * i + rn[i]] = 1.f;
    return 72;
}
/* games/limitholdem/game.py:216-231 */
static int limit_over(const void *s) {
    const limit_t *g = (const limit_t *)s;
    return ((!g->folded[0]) + (!g->folded[1]) == 1) || g->round_counter >= 4;
}
static int limit_player(const void *s) { return ((const limit_t *)s)->game_pointer; }
/* games/limitholdem/judger.py:45-85 (one pot) */
static void split_pot(const int *in_chips, const int *winners, int *allocated, int *after) {
    int nw = 0, np_ = 0;
    for (int i = 0; i < 2; i++) { nw += (winners[i] && in_chips[i] > 0); np_ += in_chips[i] > 0; }
    if (nw == 0 || nw == np_) { for (int i = 0; i < 2; i++) { allocated[i] = in_chips[i]; after[i] = 0; } return; }
    int amt = 1 << 30;
    for (int i = 0; i < 2; i++) if (in_chips[i] > 0 && in_chips[i] < amt) amt = in_chips[i];
    int one = amt * np_ / nw;   /* remainder is 0 with 2 players (judger.py:80-83 unreachable) */
    for (int i = 0; i < 2; i++) {
        allocated[i] = 0; after[i] = in_chips[i];
        if (in_chips[i] == 0) continue;
        if (winners[i]) allocated[i] += one;
        after[i] -= amt;
    }
}
/* games/limitholdem/game.py:233-243, judger.py:11-43, :87-108, utils.py:526-569 */
static void limit_payoffs(const void *s, double *out) {
    const limit_t *g = (const limit_t *)s;
    int in_chips[2] = { g->in_chips[0], g->in_chips[1] }, has_hand[2] = { !g->folded[0], !g->folded[1] };
    int payoffs[2] = {0, 0}, remaining = in_chips[0] + in_chips[1];
    uint32_t str[2] = {0, 0};
    if (has_hand[0] && has_hand[1] && g->n_public == 5)
        for (int p = 0; p < 2; p++) {
            uint8_t c[7] = { (uint8_t)g->hand[p][0], (uint8_t)g->hand[p][1] };
            for (int k = 0; k < 5; k++) c[2 + k] = (uint8_t)g->public_cards[k];
            str[p] = orc_holdem_strength7(c);
        }
    while (remaining > 0) {
        int winners[2] = {0, 0};
        if (has_hand[0] != has_hand[1]) { winners[0] = has_hand[0]; winners[1] = has_hand[1]; }   /* one folded */
        else if (!has_hand[0]) break;
        else { winners[0] = str[0] >= str[1]; winners[1] = str[1] >= str[0]; }
        int chips[2] = { in_chips[0], in_chips[1] }, each_win[2] = {0, 0}, alloc[2], after[2];
        while (chips[0] > 0 || chips[1] > 0) {            /* judger.py:87-108 */
            split_pot(chips, winners, alloc, after);
            for (int i = 0; i < 2; i++) { each_win[i] += alloc[i]; chips[i] = after[i]; }
        }
        for (int i = 0; i < 2; i++) {
            if (winners[i]) { remaining -= each_win[i]; payoffs[i] += each_win[i] - in_chips[i]; has_hand[i] = 0; in_chips[i] = 0; }
            else if (in_chips[i] > 0) { payoffs[i] += each_win[i] - in_chips[i]; in_chips[i] = each_win[i]; }
        }
    }
    out[0] = payoffs[0] / 2.0; out[1] = payoffs[1] / 2.0;
}
/* known-answer helper: limitholdem/utils.py compare_hands over P hands of 7 card ids (cards[p][0] == 255:
 * folded / None): winners = the unfolded hands of maximal strength */
void orc_holdem_winners(const uint8_t *cards, int P, uint8_t *win) {
    uint32_t best = 0, str[8];
    for (int p = 0; p < P; p++) {
        str[p] = cards[7 * p] == 255 ? 0u : 1u + orc_holdem_strength7(cards + 7 * p);
        if (str[p] > best) best = str[p];
    }
    for (int p = 0; p < P; p++) win[p] = (uint8_t)(str[p] != 0 && str[p] == best);
}
const orc_game_vt orc_vt_limit = { "limit-holdem", 2, 4, {72, 72, 0, 0}, sizeof(limit_t), limit_create, limit_reset,
    limit_step, limit_legal, limit_obs, limit_over, limit_player, limit_payoffs };
