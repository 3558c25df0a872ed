/*
 * oracle/orc_uno.c -- CPU ORACLE (test infrastructure, see orc.h).
 * UNO, 2 players, 108 cards, 61 actions: explicit ordered lists of card objects like the
 * reference (games/uno/{game,round,dealer,card,utils}.py, envs/uno.py).
 * A card is (code, color): code = 15*original_colour + trait is the frozen UnoCard.str
 * (card.py:22, Q-UNO1), color is the mutable UnoCard.color attribute.
 */
#include "orc.h"
#include <string.h>

typedef struct { uint8_t code, color; } ucard;
typedef struct {
    ucard deck[108], played[108], hand[2][108], target;
    int dl, pl, hl[2];
    int current, direction, is_over, winner, err;
    int opening;        /* throughput spec: the 15 opening cards are popped in order after a partial Fisher-Yates */
} uno_t;

#define TRAIT(c) ((c).code % 15)
static int is_wild_type(ucard c) { return TRAIT(c) >= 13; }          /* type 'wild' */
static int is_number(ucard c) { return TRAIT(c) < 10; }

/* games/uno/utils.py:31-52 */
static int uno_init_deck(ucard *d) {
    int n = 0;
    for (int col = 0; col < 4; col++) {
        for (int t = 0; t < 10; t++) { d[n].code = (uint8_t)(15 * col + t); d[n++].color = (uint8_t)col;
            if (t != 0) { d[n].code = (uint8_t)(15 * col + t); d[n++].color = (uint8_t)col; } }
        for (int t = 10; t < 13; t++) for (int k = 0; k < 2; k++) { d[n].code = (uint8_t)(15 * col + t); d[n++].color = (uint8_t)col; }
        for (int t = 13; t < 15; t++) { d[n].code = (uint8_t)(15 * col + t); d[n++].color = (uint8_t)col; }
    }
    return n;
}
/* Throughput (Philox) spec, shared with the CUDA kernels: the order of the draw pile is only
 * ever observed through pop(), so "shuffle, then pop" is replaced by "pop a uniformly random remaining
 * card": shuffles draw nothing, and a pop draws r = below(pile size) and takes the card whose code
 * (15*original colour + trait) contains position r in ascending-code order.  Cards of one code are
 * interchangeable (a wild's mutable colour is never read while it is in a pile or a hand). */
static void uno_shuffle(orc_chance *ch, ucard *x, int n) {              /* dealer.py:14-17 */
    if (ch->kind == ORC_CHANCE_PHILOX) return;
    for (int i = n - 1; i >= 1; i--) { uint32_t j = orc_below(ch, (uint32_t)i + 1u); ucard t = x[i]; x[i] = x[j]; x[j] = t; }
}
static int uno_pop(uno_t *g, orc_chance *ch, ucard *out) {              /* deck.pop(); Q-UNO4: empty -> flag, not emulated */
    if (g->dl <= 0) { g->err |= 8; return 0; }
    if (ch->kind == ORC_CHANCE_PHILOX && !g->opening) {
        int cnt[60], r = (int)orc_below(ch, (uint32_t)g->dl), code = 0;
        memset(cnt, 0, sizeof cnt);
        for (int i = 0; i < g->dl; i++) cnt[g->deck[i].code]++;
        while (r >= cnt[code]) { r -= cnt[code]; code++; }
        int at = 0;
        while (g->deck[at].code != code) at++;
        *out = g->deck[at];
        g->deck[at] = g->deck[--g->dl];
        return 1;
    }
    *out = g->deck[--g->dl];
    return 1;
}
static void uno_deal(uno_t *g, orc_chance *ch, int p, int num) {        /* dealer.py:19-26 */
    ucard c;
    for (int k = 0; k < num; k++) if (uno_pop(g, ch, &c)) g->hand[p][g->hl[p]++] = c;
}
static void uno_replace_deck(uno_t *g, orc_chance *ch) {                /* round.py:155-160 */
    for (int k = 0; k < g->pl; k++) g->deck[g->dl++] = g->played[k];
    uno_shuffle(ch, g->deck, g->dl);
    g->pl = 0;
}
static void uno_create(void *s) { (void)s; }
/* games/uno/game.py:22-56, round.py:24-52, dealer.py:28-39 */
static int uno_reset(void *s, orc_chance *ch) {
    uno_t *g = (uno_t *)s;
    memset(g, 0, sizeof *g);
    g->dl = uno_init_deck(g->deck);
    if (ch->kind == ORC_CHANCE_PHILOX) {
        /* throughput spec: only the 15 positions the opening deal pops are shuffled (reverse Fisher-Yates steps
         * i = 107 .. 93 over the deck in utils.py:31-52 order); what stays in the pile is a multiset anyway */
        for (int i = 107; i >= 93; i--) { uint32_t j = orc_below(ch, (uint32_t)i + 1u); ucard t = g->deck[i]; g->deck[i] = g->deck[j]; g->deck[j] = t; }
        g->opening = 1;
    } else uno_shuffle(ch, g->deck, g->dl);
    uno_deal(g, ch, 0, 7); uno_deal(g, ch, 1, 7);
    g->direction = 1; g->current = 0; g->winner = -1;
    ucard top;
    uno_pop(g, ch, &top);
    g->opening = 0;
    while (TRAIT(top) == 14) { g->deck[g->dl++] = top; uno_shuffle(ch, g->deck, g->dl); uno_pop(g, ch, &top); }
    if (TRAIT(top) == 13) top.color = (uint8_t)orc_below(ch, 4);
    g->target = top; g->played[g->pl++] = top;
    if (TRAIT(top) == 10) g->current = 1;
    else if (TRAIT(top) == 11) { g->direction = -1; g->current = 1; }
    else if (TRAIT(top) == 12) uno_deal(g, ch, g->current, 2);
    return g->current;
}
/* round.py:194-227 */
static void uno_non_number(uno_t *g, orc_chance *ch, ucard card) {
    int current = g->current, direction = g->direction, t = TRAIT(card);
    if (t == 11) g->direction = -direction;
    else if (t == 10) current = ((current + direction) % 2 + 2) % 2;
    else if (t == 12 || t == 14) {
        int need = t == 12 ? 2 : 4;
        if (g->dl < need) uno_replace_deck(g, ch);
        uno_deal(g, ch, ((current + direction) % 2 + 2) % 2, need);
        current = ((current + direction) % 2 + 2) % 2;
    }
    g->current = ((current + g->direction) % 2 + 2) % 2;
    g->target = card;
}
/* round.py:96-135 -> 61-bit set */
static uint64_t uno_legal_bits(const uno_t *g) {
    uint64_t legal = 0, wild4 = 0;
    const ucard *hand = g->hand[g->current]; int n = g->hl[g->current];
    ucard tg = g->target;
    for (int i = 0; i < n; i++) {
        ucard c = hand[i];
        if (is_wild_type(c)) {
            if (TRAIT(c) == 14) wild4 |= (1ull << 14) | (1ull << 29) | (1ull << 44) | (1ull << 59);
            else legal |= (1ull << 13) | (1ull << 28) | (1ull << 43) | (1ull << 58);
        } else if (c.color == tg.color || (!is_wild_type(tg) && TRAIT(c) == TRAIT(tg))) legal |= 1ull << c.code;
    }
    if (!legal) legal = wild4;
    if (!legal) legal = 1ull << 60;
    return legal;
}
/* round.py:162-192 */
static void uno_draw(uno_t *g, orc_chance *ch) {
    ucard card;
    if (g->dl == 0) uno_replace_deck(g, ch);
    if (!uno_pop(g, ch, &card)) { g->current ^= 1; return; }
    if (is_wild_type(card)) {
        card.color = (uint8_t)orc_below(ch, 4);
        g->target = card; g->played[g->pl++] = card; g->current = ((g->current + g->direction) % 2 + 2) % 2;
    } else if (card.color == g->target.color) {
        if (is_number(card)) { g->target = card; g->played[g->pl++] = card; g->current = ((g->current + g->direction) % 2 + 2) % 2; }
        else { g->played[g->pl++] = card; uno_non_number(g, ch, card); }
    } else { g->hand[g->current][g->hl[g->current]++] = card; g->current = ((g->current + g->direction) % 2 + 2) % 2; }
}
/* env.py:65-86, envs/uno.py:39-45, game.py:58-81, round.py:54-94 */
static int uno_step(void *s, orc_chance *ch, int id) {
    uno_t *g = (uno_t *)s;
    uint64_t legal = uno_legal_bits(g);
    if (id < 0 || id > 60 || !((legal >> id) & 1)) {        /* reference: np.random.choice(legal) on the GLOBAL rng; */
        g->err |= 4; id = __builtin_ctzll(legal);            /* replay uses legal ids only; fallback = lowest legal id */
    }
    if (id == 60) { uno_draw(g, ch); return g->current; }
    int p = g->current, color = id / 15, trait = id % 15, idx = -1;
    ucard *hand = g->hand[p];
    for (int i = 0; i < g->hl[p]; i++) {
        if (trait >= 13) { if (TRAIT(hand[i]) == trait) { hand[i].color = (uint8_t)color; idx = i; break; } }
        else if (hand[i].color == color && TRAIT(hand[i]) == trait && !is_wild_type(hand[i])) { idx = i; break; }
    }
    ucard card = hand[idx];
    memmove(hand + idx, hand + idx + 1, sizeof(ucard) * (size_t)(g->hl[p] - idx - 1));
    g->hl[p]--;
    if (g->hl[p] == 0) { g->is_over = 1; g->winner = p; }
    g->played[g->pl++] = card;
    if (is_number(card)) { g->current = ((g->current + g->direction) % 2 + 2) % 2; g->target = card; }
    else uno_non_number(g, ch, card);
    return g->current;
}
static int uno_legal(const void *s, uint8_t *mask) {        /* envs/uno.py:47-50 */
    uint64_t b = uno_legal_bits((const uno_t *)s); int c = 0;
    for (int a = 0; a < 61; a++) { mask[a] = (uint8_t)((b >> a) & 1); c += mask[a]; }
    return c;
}
/* envs/uno.py:24-33, games/uno/utils.py:69-127 */
static int uno_obs(const void *s, int seat, float *o) {
    const uno_t *g = (const uno_t *)s;
    if (seat < 0) seat = g->current;
    memset(o, 0, 240 * sizeof(float));
    for (int k = 0; k < 60; k++) o[k] = 1.f;                 /* plane 0 starts all ones */
    int cnt[4][15]; memset(cnt, 0, sizeof cnt);
    for (int i = 0; i < g->hl[seat]; i++) cnt[g->hand[seat][i].color][TRAIT(g->hand[seat][i])]++;   /* get_str(): current colour */
    for (int c = 0; c < 4; c++) for (int t = 0; t < 15; t++) {
        if (!cnt[c][t]) continue;
        if (t >= 13) { for (int k = 0; k < 4; k++) { o[15 * k + t] = 0.f; o[60 + 15 * k + t] = 1.f; } }
        else { o[15 * c + t] = 0.f; o[60 * cnt[c][t] + 15 * c + t] = 1.f; }
    }
    o[180 + g->target.code] = 1.f;                            /* target.str: original colour (Q-UNO1) */
    return 240;
}
static int uno_over(const void *s) { return ((const uno_t *)s)->is_over; }   /* game.py:154-160 */
static int uno_player(const void *s) { return ((const uno_t *)s)->current; }
static void uno_payoffs(const void *s, double *out) {        /* game.py:108-118 */
    const uno_t *g = (const uno_t *)s;
    out[0] = out[1] = 0;
    if (g->winner >= 0) { out[g->winner] = 1; out[1 - g->winner] = -1; }
}
/* known-answer helper: encode_hand + encode_target (games/uno/utils.py:86-127) of an arbitrary hand of card
 * codes (15*colour + trait) and target code */
int orc_uno_encode(const uint8_t *codes, int n, int target_code, float *out) {
    static uno_t g;
    memset(&g, 0, sizeof g);
    for (int i = 0; i < n; i++) { g.hand[0][i].code = codes[i]; g.hand[0][i].color = (uint8_t)(codes[i] / 15); }
    g.hl[0] = n; g.current = 0; g.winner = -1;
    g.target.code = (uint8_t)target_code; g.target.color = (uint8_t)(target_code / 15);
    return uno_obs(&g, 0, out);
}
const orc_game_vt orc_vt_uno = { "uno", 2, 61, {240, 240, 0, 0}, sizeof(uno_t), uno_create, uno_reset, uno_step,
    uno_legal, uno_obs, uno_over, uno_player, uno_payoffs };
