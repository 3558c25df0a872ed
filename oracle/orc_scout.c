/*
 * oracle/orc_scout.c -- CPU ORACLE (test infrastructure, see orc.h).
 * Scout, 4 players, 45 cards, 204 actions, explicit ordered lists like the reference
 * (games/scout/{game,round,dealer,player,card,judger}.py, games/scout/utils/utils.py, envs/scout.py).
 */
#include "orc.h"
#include <string.h>

typedef struct { uint8_t top, bottom; } scard;
typedef struct {
    scard deck[45], hand[4][24], table[24];
    int dl, hl[4], tl, score[4], owner, consecutive, game_over, cur, forced;
} scout_t;

static void sc_shuffle(orc_chance *ch, scard *x, int n) {               /* dealer.py:19-22 */
    for (int i = n - 1; i >= 1; i--) { uint32_t j = orc_below(ch, (uint32_t)i + 1u); scard t = x[i]; x[i] = x[j]; x[j] = t; }
}
/* utils/utils.py:17-67 */
static int sc_valid_segment(const scard *c, int len) {
    if (len < 2) return 1;
    int group = 1, asc = 1, desc = 1;
    for (int i = 0; i < len; i++) if (c[i].top != c[0].top) group = 0;
    if (group) return 1;
    for (int i = 0; i + 1 < len; i++) {
        int d = (int)c[i + 1].top - (int)c[i].top;
        if (d != 1) asc = 0;
        if (d != -1) desc = 0;
        if (!asc && !desc) return 0;
    }
    return asc || desc;
}
/* utils/utils.py:144-184 */
static void sc_strength(const scard *c, int len, int *type, int *rank) {
    if (len == 1) { *type = 0; *rank = c[0].top; return; }
    int group = 1;
    for (int i = 0; i < len; i++) if (c[i].top != c[0].top) group = 0;
    if (group) { *type = 2; *rank = c[0].top; return; }
    int asc = 1, desc = 1;
    for (int i = 0; i + 1 < len; i++) {
        int d = (int)c[i + 1].top - (int)c[i].top;
        if (d != 1) asc = 0;
        if (d != -1) desc = 0;
        if (!asc && !desc) break;
    }
    if (asc || desc) { *type = 1; *rank = c[0].top > c[len - 1].top ? c[0].top : c[len - 1].top; return; }
    *type = 0; *rank = 0;
    for (int i = 0; i < len; i++) if (c[i].top > *rank) *rank = c[i].top;
}
/* round.py:262-295 */
static int sc_stronger(const scard *a, int alen, const scard *b, int blen) {
    if (blen == 0) return 1;
    if (alen > blen) return 1;
    if (alen < blen) return 0;
    int ta, ra, tb, rb;
    sc_strength(a, alen, &ta, &ra); sc_strength(b, blen, &tb, &rb);
    if (ta > tb) return 1;
    if (ta < tb) return 0;
    return ra > rb;
}
static int sc_play_id(int s, int e) { int id = 0; for (int k = 0; k < s; k++) id += 16 - k; return id + e - s - 1; }   /* utils.py:186-200 */
/* round.py:225-260 -> 204-entry mask; sets the forced-scout flag */
static int sc_legal(scout_t *g, uint8_t *mask) {
    const scard *h = g->hand[g->cur]; int n = g->hl[g->cur], cnt = 0, playable = 0;
    memset(mask, 0, 204);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j <= n; j++)
            if (sc_valid_segment(h + i, j - i) && (g->tl == 0 || sc_stronger(h + i, j - i, g->table, g->tl))) {
                if (j <= 16 && i < 16) { mask[sc_play_id(i, j)] = 1; cnt++; }
                playable++;
            }
    g->forced = playable == 0;
    if (g->tl > 0 && n < 16)
        for (int ins = 0; ins <= n; ins++) {
            mask[136 + 4 * ins] = mask[136 + 4 * ins + 1] = 1; cnt += 2;
            if (g->tl > 1) { mask[136 + 4 * ins + 2] = mask[136 + 4 * ins + 3] = 1; cnt += 2; }
        }
    return cnt;
}
static void sc_create(void *s) { (void)s; }
/* games/scout/game.py:37-65, dealer.py:12-17, round.py:22-50 (second shuffle, Q-SC1) */
static int sc_reset(void *s, orc_chance *ch) {
    scout_t *g = (scout_t *)s; uint8_t m[204];
    memset(g, 0, sizeof *g);
    for (int top = 1; top <= 10; top++) for (int bot = top + 1; bot <= 10; bot++) { g->deck[g->dl].top = (uint8_t)top; g->deck[g->dl++].bottom = (uint8_t)bot; }
    sc_shuffle(ch, g->deck, 45);
    sc_shuffle(ch, g->deck, 45);
    for (int p = 0; g->dl > 0; p = (p + 1) % 4) g->hand[p][g->hl[p]++] = g->deck[--g->dl];
    g->owner = -1;
    g->cur = (int)orc_below(ch, 4);
    sc_legal(g, m);
    return g->cur;
}
/* envs/scout.py:53-80,130-133; round.py:63-153 */
static int sc_step(void *s, orc_chance *ch, int id) {
    scout_t *g = (scout_t *)s; (void)ch;
    uint8_t m[204];
    sc_legal(g, m);
    if (id < 0 || id >= 204 || !m[id]) { for (id = 0; id < 204 && !m[id]; id++) {} }   /* reference raises; replay uses legal ids only */
    int p = g->cur, forced = g->forced;
    if (id < 136) {
        int st = 0, k = id;
        while (k >= 16 - st) { k -= 16 - st; st++; }
        int en = st + k + 1, len = en - st;
        scard seg[24];
        memcpy(seg, g->hand[p] + st, sizeof(scard) * (size_t)len);
        memmove(g->hand[p] + st, g->hand[p] + en, sizeof(scard) * (size_t)(g->hl[p] - en));
        g->hl[p] -= len;
        if (g->tl) g->score[p] += g->tl;
        memcpy(g->table, seg, sizeof(scard) * (size_t)len); g->tl = len;
        g->owner = p; g->consecutive = 0;
    } else {
        int k = id - 136, ins = k / 4, front = (k % 4) < 2, flip = k % 2;
        scard c;
        if (front) { c = g->table[0]; memmove(g->table, g->table + 1, sizeof(scard) * (size_t)(g->tl - 1)); }
        else c = g->table[g->tl - 1];
        g->tl--;
        if (flip) { uint8_t t = c.top; c.top = c.bottom; c.bottom = t; }
        if (ins > g->hl[p]) ins = g->hl[p];                                     /* list.insert clamps */
        memmove(g->hand[p] + ins + 1, g->hand[p] + ins, sizeof(scard) * (size_t)(g->hl[p] - ins));
        g->hand[p][ins] = c; g->hl[p]++;
        if (forced && g->owner >= 0) g->score[g->owner] += 1;                   /* Q-SC2 */
        g->consecutive++;
        if (g->tl == 0) { g->owner = -1; g->consecutive = 0; }
        if (g->consecutive == 3 && g->owner >= 0) g->game_over = 1;
    }
    g->cur = (p + 1) % 4;
    if (sc_legal(g, m) == 0) g->game_over = 1;
    return g->cur;
}
static int sc_legal_c(const void *s, uint8_t *mask) { scout_t tmp = *(const scout_t *)s; return sc_legal(&tmp, mask); }
/* envs/scout.py:140-248 */
static int sc_obs(const void *s, int seat, float *o) {
    const scout_t *g = (const scout_t *)s;
    if (seat < 0) seat = g->cur;
    memset(o, 0, 688 * sizeof(float));
    for (int i = 0; i < g->hl[seat] && i < 16; i++) {
        o[i * 10 + g->hand[seat][i].top - 1] = 1.f; o[160 + i * 10 + g->hand[seat][i].bottom - 1] = 1.f; o[640 + i] = 1.f;
    }
    for (int j = 0; j < g->tl && j < 16; j++) {
        o[320 + j * 10 + g->table[j].top - 1] = 1.f; o[480 + j * 10 + g->table[j].bottom - 1] = 1.f; o[656 + j] = 1.f;
    }
    float *info = o + 672;
    info[g->owner >= 0 ? g->owner : 4] = 1.f;
    info[5] = (float)((double)g->consecutive / 3.0);
    for (int p = 0; p < 4; p++) info[6 + p] = (float)((double)g->hl[p] / 16.0);
    info[10] = (float)((double)g->score[seat] / 16.0);
    info[11] = (float)((double)g->tl / 16.0);
    info[12] = 0.f;                                                             /* orientation never flipped in automated play */
    info[13] = g->forced ? 1.f : 0.f;                                           /* Q-SC3: current player's flag */
    info[14] = g->tl ? (float)((double)g->table[0].top / 10.0) : 0.f;
    info[15] = g->tl ? (float)((double)g->table[g->tl - 1].top / 10.0) : 0.f;
    return 688;
}
static int sc_over(const void *s) { return ((const scout_t *)s)->game_over; }
static int sc_player(const void *s) { return ((const scout_t *)s)->cur; }
static void sc_payoffs(const void *s, double *out) {                            /* judger.py:15-30 */
    const scout_t *g = (const scout_t *)s;
    for (int p = 0; p < 4; p++) out[p] = g->score[p] - g->hl[p];
}
const orc_game_vt orc_vt_scout = { "scout", 4, 204, {688, 688, 688, 688}, sizeof(scout_t), sc_create, sc_reset, sc_step,
    sc_legal_c, sc_obs, sc_over, sc_player, sc_payoffs };
