// game_poker.cuh -- Leduc Hold'em and Limit Texas Hold'em (2 players) as register-resident
// state machines, one env per thread.  Reference: rlcard/games/leducholdem/*, rlcard/games/
// limitholdem/* (round.py is shared), rlcard/envs/{leducholdem,limitholdem}.py.
#pragma once
#include "common.cuh"

namespace rlc {

enum { kCall = 0, kRaise = 1, kFold = 2, kCheck = 3 };   // envs/leducholdem.py:26

// Reverse Fisher-Yates (RandomState.shuffle) over an identity deck of n cards of which only the last K
// positions are ever popped: step s swaps positions n-1-s and j[s]; c[s] is the card left at position
// n-1-s, i.e. the s-th deck.pop().  The card is found by tracing position j[s] back through the earlier
// swaps (it can only have been moved by a swap whose partner j[r] it equals) -- no deck array needed.
template <int K>
__device__ __forceinline__ void fy_tail_cards(int n, const int (&j)[K], int (&c)[K]) {
#pragma unroll
    for (int s = 0; s < K; s++) {
        int pos = j[s];
#pragma unroll
        for (int r = s - 1; r >= 0; r--) pos = (pos == j[r]) ? n - 1 - r : pos;
        c[s] = pos;
    }
}

// games/limitholdem/round.py as plain registers (2 players)
struct BetRound {
    int pointer, have_raised, not_raise_num, raised0, raised1;
    __device__ __forceinline__ void start(int p, int r0, int r1) {          // round.py:35-51
        pointer = p; have_raised = 0; not_raise_num = 0; raised0 = r0; raised1 = r1;
    }
    __device__ __forceinline__ uint32_t legal(int allowed) const {          // round.py:95-116
        const int mx = max(raised0, raised1), mine = pointer ? raised1 : raised0;
        uint32_t m = 0xFu;
        if (have_raised >= allowed) m &= ~(1u << kRaise);
        if (mine < mx) m &= ~(1u << kCheck);
        if (mine == mx) m &= ~(1u << kCall);
        return m;
    }
    // round.py:53-93; returns chips the actor adds, sets folded flag of the actor
    __device__ __forceinline__ int proceed(int action, int raise_amount, bool &fold) {   // branch free (lanes act differently)
        const int mx = max(raised0, raised1), mine = pointer ? raised1 : raised0;
        const bool call = action == kCall, raise = action == kRaise;
        fold = action == kFold;
        const int diff = call ? mx - mine : (raise ? mx - mine + raise_amount : 0);
        have_raised += raise ? 1 : 0;
        not_raise_num = raise ? 1 : (fold ? not_raise_num : not_raise_num + 1);
        raised1 += pointer ? diff : 0; raised0 += pointer ? 0 : diff;
        pointer ^= 1;     // 2 players: the other seat is never the folded one
        return diff;
    }
};

// ========================================================================================
// Leduc: 32-bit packed state.
//  [0:2) hand0 rank  [2:4) hand1 rank  [4:6) public rank (pre-drawn third pop)  [6] public dealt
//  [7:11) chips0 [11:15) chips1 [15:19) raised0 [19:23) raised1 [23:25) have_raised
//  [25:27) not_raise_num [27] pointer [28:30) round_counter [30] folded0 [31] folded1
// ========================================================================================
struct Leduc {
    static constexpr int kGameId = 1, P = 2, A = 4, OBS = 36, GAME_WORDS = 1, MASK_WORDS = 1;
    static constexpr bool kUsesChain = true;   // reset draws ride on the policy word (common.cuh chain())
    static constexpr int kMaxResetDraws = 6;
    static constexpr int kRolloutMinEpw = 32;    // measured: fewer envs per warp only adds idle lanes (kernels.cuh)
    static constexpr bool kHasApply = false;
    static constexpr bool kWarpDeal = false;
    static constexpr bool kChanceAwareState = false;
    static constexpr int kSharedBytes = 128;   // deal table: x in [0,120) -> hand0 | hand1 << 2 | public << 4
    int hand0, hand1, pub, pub_dealt, chips0, chips1, rc, fold0, fold1;
    BetRound r;
    const uint8_t *deal_lut;

    // the 120 ordered draws of 3 cards out of [SJ,HJ,SQ,HQ,SK,HK] as ranks, by Fisher-Yates code
    static __device__ __forceinline__ uint32_t deal_code(uint32_t x) {
        const int j[3] = { (int)(x / 20u), (int)((x >> 2) % 5u), (int)(x & 3u) };
        int c[3];
        fy_tail_cards<3>(6, j, c);                           // cards popped from positions 5,4,3; rank = card >> 1
        return (uint32_t)((c[0] >> 1) | ((c[1] >> 1) << 2) | ((c[2] >> 1) << 4));
    }
    static __device__ __forceinline__ void fill_shared(uint8_t *sm, int tid, int nthreads) {
        for (int x = tid; x < 120; x += nthreads) sm[x] = (uint8_t)deal_code((uint32_t)x);
    }
    __device__ __forceinline__ void bind_shared(const uint8_t *sm) { deal_lut = sm; }

    __device__ __forceinline__ void load(const uint32_t *st, size_t n, size_t i) {
        const uint32_t w = st[i];
        hand0 = bf_get(w, 0, 2); hand1 = bf_get(w, 2, 2); pub = bf_get(w, 4, 2); pub_dealt = bf_get(w, 6, 1);
        chips0 = bf_get(w, 7, 4); chips1 = bf_get(w, 11, 4); r.raised0 = bf_get(w, 15, 4); r.raised1 = bf_get(w, 19, 4);
        r.have_raised = bf_get(w, 23, 2); r.not_raise_num = bf_get(w, 25, 2); r.pointer = bf_get(w, 27, 1);
        rc = bf_get(w, 28, 2); fold0 = bf_get(w, 30, 1); fold1 = bf_get(w, 31, 1);
        (void)n;
    }
    __device__ __forceinline__ void store(uint32_t *st, size_t n, size_t i) const {
        st[i] = hand0 | (hand1 << 2) | (pub << 4) | (pub_dealt << 6) | (chips0 << 7) | (chips1 << 11) |
                (r.raised0 << 15) | (r.raised1 << 19) | (r.have_raised << 23) | (r.not_raise_num << 25) |
                (r.pointer << 27) | (rc << 28) | (fold0 << 30) | ((uint32_t)fold1 << 31);
        (void)n;
    }
    // games/leducholdem/game.py:46-95; deck [SJ,HJ,SQ,HQ,SK,HK] (dealer.py:10) kept as nibbles
    template <class Ch> __device__ __forceinline__ void reset(Ch &ch) {
        uint32_t code;
        if constexpr (Ch::kKind == 0) {                      // throughput: the three cards ride on the policy word
            code = deal_lut[ch.chain(120u)];                 // x = 20 j0 + 4 j1 + j2: swap partners of positions 5,4,3
        } else {                                             // replay: the reference's draws, i = 5,4,3 (2,1 unobserved)
            const uint32_t j0 = ch.below(6u), j1 = ch.below(5u), j2 = ch.below(4u);
            ch.skip_fy(2, 1);
            code = deal_code(20u * j0 + 4u * j1 + j2);
        }
        hand0 = code & 3; hand1 = (code >> 2) & 3; pub = (code >> 4) & 3; pub_dealt = 0;
        const int sb = (int)ch.chain(2u);
        chips0 = sb == 0 ? 1 : 2; chips1 = sb == 0 ? 2 : 1;
        fold0 = fold1 = 0; rc = 0;
        r.start(sb, chips0, chips1);
    }
    __device__ __forceinline__ int raise_amount() const { return rc == 0 ? 2 : 4; }   // game.py:36,129
    __device__ __forceinline__ void legal(uint32_t (&m)[1]) const { m[0] = r.legal(2); }
    __device__ __forceinline__ int player() const { return r.pointer; }
    __device__ __forceinline__ bool over() const { return (fold0 + fold1 == 1) || rc >= 2; }   // game.py:154-168
    // env.py:65-86, envs/leducholdem.py:81-96, game.py:97-136
    template <class Ch> __device__ __forceinline__ void step(int id, Ch &, int &err) {
        const uint32_t m = r.legal(2);
        int action = id;
        if (id < 0 || id > 3 || !((m >> id) & 1u)) { action = ((m >> kCheck) & 1u) ? kCheck : kFold; err |= 4; }
        const int actor = r.pointer;
        bool fold;
        const int diff = r.proceed(action, raise_amount(), fold);
        if (actor) { chips1 += diff; fold1 |= fold; } else { chips0 += diff; fold0 |= fold; }
        if (r.not_raise_num >= 2) {
            if (rc == 0) pub_dealt = 1;
            rc++;
            r.start(r.pointer, 0, 0);
        }
    }
    // games/leducholdem/judger.py:12-64, game.py:170-178 (chips / big blind)
    __device__ __forceinline__ void payoffs(float *out) const {
        int w0, w1;
        if (fold0 + fold1 == 1) { w0 = fold1; w1 = fold0; }
        else if (pub_dealt && hand0 == pub) { w0 = 1; w1 = 0; }
        else if (pub_dealt && hand1 == pub) { w0 = 0; w1 = 1; }
        else { w0 = hand0 >= hand1; w1 = hand1 >= hand0; }
        // winner(s) share the pot, minus own chips, / big blind 2 -- in exact quarter-chip integers:
        // split: ((c0+c1)/2 - c0)/2 = (c1-c0)/4; sole winner: c_other/2; loser: -c_own/2
        const int tie = w0 & w1;
        const int q0 = tie ? chips1 - chips0 : (w0 ? 2 * chips1 : -2 * chips0);
        const int q1 = tie ? chips0 - chips1 : (w1 ? 2 * chips0 : -2 * chips1);
        out[0] = (float)q0 * 0.25f; out[1] = (float)q1 * 0.25f;
    }
    // envs/leducholdem.py:41-71 (row is pre-zeroed)
    template <class T> __device__ __forceinline__ void encode_obs(int seat, bool, T *row) const {
        const int mine = seat ? chips1 : chips0, other = seat ? chips0 : chips1;
        row[seat ? hand1 : hand0] = (T)1;
        if (pub_dealt) row[3 + pub] = (T)1;
        row[6 + mine] = (T)1;
        row[21 + other] = (T)1;
    }
};

// ========================================================================================
// 7-card evaluator on rank/suit bit masks (orders hands like games/limitholdem/utils.py
// compare_hands).  card id = 13*suit + rank (A=0, 2..K = 1..12).
// ========================================================================================
// keep the `keep` highest set bits.  Branch free: seven distinct cards leave at most keep + 3 ranks in any mask this is
// called on (the kickers of a made hand), so three predicated removals of the lowest bit suffice.
__device__ __forceinline__ uint32_t top_bits(uint32_t m, int keep) {
    const int r = __popc(m) - keep;
    m = r > 0 ? m & (m - 1) : m;
    m = r > 1 ? m & (m - 1) : m;
    m = r > 2 ? m & (m - 1) : m;
    return m;
}
__device__ __forceinline__ uint32_t top_bit(uint32_t m) { return 31u - (uint32_t)__clz(m | 1u); }   // 0 for m == 0
__device__ __forceinline__ int straight_top(uint32_t m) {   // m bit r = rank (2->0 .. A->12); 0 = none, else top+1
    const uint32_t x = (m << 1) | ((m >> 12) & 1u);          // bit0 = ace playing low
    const uint32_t run = x & (x >> 1) & (x >> 2) & (x >> 3) & (x >> 4);
    return run ? (32 - __clz(run)) + 3 : 0;                  // top bit index in x (1-based rank), 5-high wheel -> 4
}
// rank bit (2 -> bit 0 .. A -> bit 12) of card id c = 13*suit + rank (A = 0, 2..K = 1..12) into the suit's mask
__device__ __forceinline__ void holdem_add_card(uint32_t (&s)[4], int c) {
    const int suit = c / 13, rk = c - 13 * suit;
    const uint32_t bit = 1u << (rk == 0 ? 12 : rk - 1);
    s[0] |= suit == 0 ? bit : 0u; s[1] |= suit == 1 ? bit : 0u;
    s[2] |= suit == 2 ? bit : 0u; s[3] |= suit == 3 ? bit : 0u;
}
__device__ __forceinline__ uint32_t holdem_strength_masks(const uint32_t (&s)[4]);
__device__ __forceinline__ uint32_t holdem_strength7(const int (&c)[7]) {
    uint32_t s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 7; i++) holdem_add_card(s, c[i]);
    return holdem_strength_masks(s);
}
// strength of the seven cards given as four per-suit rank masks.  Branch free: the 32 lanes of a warp hold hands of
// different categories, so every category is scored and the best valid one selected (category in bits 26.., then the
// tie-break ranks in the order games/limitholdem/utils.py compares them).
__device__ __forceinline__ uint32_t holdem_strength_masks(const uint32_t (&s)[4]) {
    const uint32_t any = s[0] | s[1] | s[2] | s[3];
    uint32_t fl = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) fl = __popc(s[k]) >= 5 ? s[k] : fl;
    const uint32_t x4 = s[0] & s[1] & s[2] & s[3];
    const uint32_t x3 = ((s[0] & s[1] & s[2]) | (s[0] & s[1] & s[3]) | (s[0] & s[2] & s[3]) | (s[1] & s[2] & s[3])) & ~x4;
    const uint32_t x2 = ((s[0] & s[1]) | (s[0] & s[2]) | (s[0] & s[3]) | (s[1] & s[2]) | (s[1] & s[3]) | (s[2] & s[3])) & ~x3 & ~x4;
    const uint32_t sf = (uint32_t)straight_top(fl), st = (uint32_t)straight_top(any);
    const uint32_t q = top_bit(x4), t3 = top_bit(x3), p2 = top_bit(x2), p2b = top_bit(x2 & ~(1u << p2));
    const uint32_t rest3 = (x3 & ~(1u << t3)) | x2;                         // what can fill a full house
    const uint32_t v9 = (9u << 26) | sf;
    const uint32_t v8 = (8u << 26) | (q << 4) | top_bit(any & ~(1u << q));
    const uint32_t v7 = (7u << 26) | (t3 << 4) | top_bit(rest3);
    const uint32_t v6 = (6u << 26) | top_bits(fl, 5);
    const uint32_t v5 = (5u << 26) | st;
    const uint32_t v4 = (4u << 26) | (t3 << 13) | top_bits(any & ~(1u << t3), 2);
    const uint32_t v3 = (3u << 26) | (p2 << 17) | (p2b << 13) | top_bits(any & ~(1u << p2) & ~(1u << p2b), 1);
    const uint32_t v2 = (2u << 26) | (p2 << 13) | top_bits(any & ~(1u << p2), 3);
    uint32_t v = (1u << 26) | top_bits(any, 5);
    v = x2 ? v2 : v;
    v = (x2 & (x2 - 1)) ? v3 : v;
    v = x3 ? v4 : v;
    v = st ? v5 : v;
    v = fl ? v6 : v;
    v = (x3 && rest3) ? v7 : v;
    v = x4 ? v8 : v;
    v = sf ? v9 : v;
    return v;
}

// ---- the same evaluator over lookup tables (north_star: "7-card hand evaluator via shared-memory lookup tables") ----------
// Two tables, tabulated ONCE per device from the functions above (tu_limit.cu k_holdem_build_tables: table == engine) and
// staged in shared memory by the kernels that score showdowns in their inner loop:
//   card64[52]  card id -> its bit in a 4 x 16-bit layout (suit s, rank bit r -> bit 16 s + r): seven cards are seven
//               64-bit loads OR-ed together instead of seven div / mod / four-way select sequences
//   t5[8192]    13-bit rank mask -> 0x8000 | straight_top(m) when m holds a straight, else its five highest bits
//               (top_bits(m, 5)): one look-up answers "straight?" / "straight flush?" and yields the flush / high-card
//               kickers
// With the tables the categories are tested in order of rank and only the winning one is scored: at a Limit showdown one
// or two lanes of the warp are active, so a branchy evaluator costs its own path (~60 instructions), not the sum of all
// nine categories (~180 branch free).  Returns exactly holdem_strength_masks().
struct HoldemLut { const unsigned long long *card64; const uint16_t *t5; };
constexpr int kHoldemLutBytes = 52 * 8 + 8192 * 2;           // card64 | t5, in this order (16-byte aligned sizes)
__device__ __forceinline__ uint32_t holdem_strength_lut(const uint32_t (&s)[4], const uint16_t *t5) {
    const uint32_t any = s[0] | s[1] | s[2] | s[3];
    uint32_t fl = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) fl = __popc(s[k]) >= 5 ? s[k] : fl;
    if (fl) {                                                  // seven cards: a flush excludes quads and full houses
        const uint32_t e = t5[fl];
        return (e & 0x8000u) ? (9u << 26) | (e & 15u) : (6u << 26) | e;
    }
    const uint32_t x4 = s[0] & s[1] & s[2] & s[3];
    if (x4) {
        const uint32_t q = top_bit(x4);
        return (8u << 26) | (q << 4) | top_bit(any & ~(1u << q));
    }
    const uint32_t x3 = (s[0] & s[1] & s[2]) | (s[0] & s[1] & s[3]) | (s[0] & s[2] & s[3]) | (s[1] & s[2] & s[3]);
    const uint32_t x2 = ((s[0] & s[1]) | (s[0] & s[2]) | (s[0] & s[3]) | (s[1] & s[2]) | (s[1] & s[3]) | (s[2] & s[3])) & ~x3;
    const uint32_t t3 = top_bit(x3);
    if (x3) {
        const uint32_t rest3 = (x3 & ~(1u << t3)) | x2;
        if (rest3) return (7u << 26) | (t3 << 4) | top_bit(rest3);
    }
    const uint32_t e = t5[any];
    if (e & 0x8000u) return (5u << 26) | (e & 15u);
    if (x3) return (4u << 26) | (t3 << 13) | top_bits(any & ~(1u << t3), 2);
    if (x2) {
        const uint32_t p2 = top_bit(x2), r2 = x2 & ~(1u << p2);
        if (r2) {
            const uint32_t p2b = top_bit(r2);
            return (3u << 26) | (p2 << 17) | (p2b << 13) | top_bits(any & ~(1u << p2) & ~(1u << p2b), 1);
        }
        return (2u << 26) | (p2 << 13) | top_bits(any & ~(1u << p2), 3);
    }
    return (1u << 26) | e;                                    // seven distinct ranks, no straight: e = the five highest
}
// who wins a showdown of the nine cards (w0 / w1 card fields of the Limit state words): 0 seat 0, 1 seat 1, 2 split
__device__ __forceinline__ int holdem_showdown_lut(uint32_t c_lo, uint32_t c_hi, const HoldemLut &lut) {
    const unsigned long long bd = lut.card64[(c_lo >> 24) & 63u] | lut.card64[c_hi & 63u] | lut.card64[(c_hi >> 6) & 63u] |
                                  lut.card64[(c_hi >> 12) & 63u] | lut.card64[(c_hi >> 18) & 63u];
    const unsigned long long h0 = bd | lut.card64[c_lo & 63u] | lut.card64[(c_lo >> 12) & 63u];
    const unsigned long long h1 = bd | lut.card64[(c_lo >> 6) & 63u] | lut.card64[(c_lo >> 18) & 63u];
    const uint32_t m0[4] = { (uint32_t)h0 & 0x1fffu, (uint32_t)(h0 >> 16) & 0x1fffu, (uint32_t)(h0 >> 32) & 0x1fffu, (uint32_t)(h0 >> 48) & 0x1fffu };
    const uint32_t m1[4] = { (uint32_t)h1 & 0x1fffu, (uint32_t)(h1 >> 16) & 0x1fffu, (uint32_t)(h1 >> 32) & 0x1fffu, (uint32_t)(h1 >> 48) & 0x1fffu };
    const uint32_t s0 = holdem_strength_lut(m0, lut.t5), s1 = holdem_strength_lut(m1, lut.t5);
    return s0 > s1 ? 0 : (s1 > s0 ? 1 : 2);
}

// ========================================================================================
// Limit Hold'em: 4 game words.
//  w0: cards 0..4 (6 bits each): hole p0a, p1a, p0b, p1b (deal order game.py:68-69), flop0
//  w1: flop1, flop2, turn, river (6 bits each) | have_raised [24:27) | not_raise_num [27:29) |
//      pointer [29] | folded0 [30] | folded1 [31]
//  w2: chips0 [0:6) chips1 [6:12) raised0 [12:17) raised1 [17:22) round_counter [22:25)
//  w3: raise_nums 4x3 bits [0:12) | raise_nums shown by the reset() state (Q-LH1) [12:24)
// ========================================================================================
struct Limit {
    static constexpr int kGameId = 2, P = 2, A = 4, OBS = 72, GAME_WORDS = 4, MASK_WORDS = 1;
    static constexpr bool kUsesChain = false;   // reset draws ride on the policy word (common.cuh chain())
    static constexpr bool kEpisodeDeal = true;  // throughput-mode deals are keyed by the episode ordinal (common.cuh deal words)
    static constexpr int kMaxResetDraws = 52;
    static constexpr int kRolloutMinEpw = 32;    // measured: fewer envs per warp only adds idle lanes (kernels.cuh)
    static constexpr bool kHasApply = false;
    static constexpr bool kWarpDeal = false;
    static constexpr bool kChanceAwareState = false;
    static constexpr int kSharedBytes = 0;
    static __device__ __forceinline__ void fill_shared(uint8_t *, int, int) {}
    __device__ __forceinline__ void bind_shared(const uint8_t *) {}
    int card[9];           // 0,2 = p0 hole; 1,3 = p1 hole; 4..8 board in deal order
    int chips0, chips1, rc, fold0, fold1;
    uint32_t rn, rn_shown;  // 4 x 3-bit raise counters
    BetRound r;

    __device__ __forceinline__ void load(const uint32_t *st, size_t n, size_t i) {
        const uint32_t w0 = st[i], w1 = st[n + i], w2 = st[2 * n + i], w3 = st[3 * n + i];
#pragma unroll
        for (int k = 0; k < 5; k++) card[k] = bf_get(w0, 6 * k, 6);
#pragma unroll
        for (int k = 0; k < 4; k++) card[5 + k] = bf_get(w1, 6 * k, 6);
        r.have_raised = bf_get(w1, 24, 3); r.not_raise_num = bf_get(w1, 27, 2); r.pointer = bf_get(w1, 29, 1);
        fold0 = bf_get(w1, 30, 1); fold1 = bf_get(w1, 31, 1);
        chips0 = bf_get(w2, 0, 6); chips1 = bf_get(w2, 6, 6); r.raised0 = bf_get(w2, 12, 5); r.raised1 = bf_get(w2, 17, 5);
        rc = bf_get(w2, 22, 3);
        rn = w3 & 0xfffu; rn_shown = (w3 >> 12) & 0xfffu;
    }
    __device__ __forceinline__ void store(uint32_t *st, size_t n, size_t i) const {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) w0 |= (uint32_t)card[k] << (6 * k);
#pragma unroll
        for (int k = 0; k < 4; k++) w1 |= (uint32_t)card[5 + k] << (6 * k);
        w1 |= (r.have_raised << 24) | (r.not_raise_num << 27) | (r.pointer << 29) | (fold0 << 30) | ((uint32_t)fold1 << 31);
        st[i] = w0; st[n + i] = w1;
        st[2 * n + i] = chips0 | (chips1 << 6) | (r.raised0 << 12) | (r.raised1 << 17) | (rc << 22);
        st[3 * n + i] = rn | (rn_shown << 12);
    }
    // games/limitholdem/game.py:46-103: only the last 9 deck positions are ever popped
    template <class Ch> __device__ __forceinline__ void reset(Ch &ch) {
        int j[9], sb;                                          // swap partners of positions 51..43
        if constexpr (Ch::kKind == 0) {                        // throughput: three deal words (keyed by the episode ordinal,
            uint32_t d0, d1, d2, d3;                           // common.cuh) give nine cards + blind
            ch.deal_words(d0, d1, d2, d3);
            const uint32_t x1 = __umulhi(d0, 52u * 51u * 50u), x2 = __umulhi(d1, 49u * 48u * 47u), x3 = __umulhi(d2, 46u * 45u * 44u * 2u);
            // mixed-radix digits, two divisions per word: x / (a * b) == (x / b) / a
            const uint32_t q1 = x1 / 50u, q2 = x2 / 47u, z = x3 >> 1, q3 = z / 44u;
            j[2] = (int)(x1 - 50u * q1); j[0] = (int)(q1 / 51u); j[1] = (int)(q1 - 51u * (uint32_t)j[0]);
            j[5] = (int)(x2 - 47u * q2); j[3] = (int)(q2 / 48u); j[4] = (int)(q2 - 48u * (uint32_t)j[3]);
            j[8] = (int)(z - 44u * q3); j[6] = (int)(q3 / 45u); j[7] = (int)(q3 - 45u * (uint32_t)j[6]);
            sb = (int)(x3 & 1u);
        } else {                                               // replay: the reference's 51 shuffle draws + randint
#pragma unroll
            for (int s = 0; s < 9; s++) j[s] = (int)ch.below((uint32_t)(52 - s));
            ch.skip_fy(42, 1);
            sb = (int)ch.below(2u);
        }
        fy_tail_cards<9>(52, j, card);                         // utils.py:34-43 deck order == card2index.json ids
        chips0 = sb == 0 ? 1 : 2; chips1 = sb == 0 ? 2 : 1;
        fold0 = fold1 = 0; rc = 0;
        r.start(sb, chips0, chips1);                          // (bb+1)%2 == sb for two players (game.py:81)
        rn_shown = rn;                                        // Q-LH1: reset() state shows the old list (game.py:98 vs :101)
        rn = 0;
    }
    __device__ __forceinline__ int n_public() const { return rc == 0 ? 0 : (rc == 1 ? 3 : (rc == 2 ? 4 : 5)); }
    __device__ __forceinline__ void legal(uint32_t (&m)[1]) const { m[0] = r.legal(4); }
    __device__ __forceinline__ int player() const { return r.pointer; }
    __device__ __forceinline__ bool over() const { return (fold0 + fold1 == 1) || rc >= 4; }   // game.py:216-231
    // game.py:105-156
    template <class Ch> __device__ __forceinline__ void step(int id, Ch &, int &err) {
        const uint32_t m = r.legal(4);
        int action = id;
        if (id < 0 || id > 3 || !((m >> id) & 1u)) { action = ((m >> kCheck) & 1u) ? kCheck : kFold; err |= 4; }
        const int actor = r.pointer;
        bool fold;
        const int diff = r.proceed(action, rc >= 2 ? 4 : 2, fold);   // doubled from round 2 on (game.py:148-149)
        if (actor) { chips1 += diff; fold1 |= fold; } else { chips0 += diff; fold0 |= fold; }
        rn = (rn & ~(7u << (3 * rc))) | ((uint32_t)r.have_raised << (3 * rc));   // game.py:133
        if (r.not_raise_num >= 2) { rc++; r.start(r.pointer, 0, 0); }
    }
    // who wins a showdown of this deal: 0 seat 0, 1 seat 1, 2 split (judger.py:11-108 over utils.py:526-614)
    __device__ __forceinline__ int showdown_outcome() const {
        uint32_t bd[4] = {0, 0, 0, 0};                               // the five board cards are shared by both hands
#pragma unroll
        for (int k = 4; k < 9; k++) holdem_add_card(bd, card[k]);
        uint32_t m0[4] = { bd[0], bd[1], bd[2], bd[3] }, m1[4] = { bd[0], bd[1], bd[2], bd[3] };
        holdem_add_card(m0, card[0]); holdem_add_card(m0, card[2]);
        holdem_add_card(m1, card[1]); holdem_add_card(m1, card[3]);
        const uint32_t s0 = holdem_strength_masks(m0), s1 = holdem_strength_masks(m1);
        return s0 > s1 ? 0 : (s1 > s0 ? 1 : 2);
    }
    // game.py:233-243, judger.py:11-108 for two players: winner takes min(chips) from the loser
    __device__ __forceinline__ void payoffs(float *out) const {
        // random Limit play mostly ends by a fold (a showdown needs eight actions without one), so the evaluator stays
        // behind a branch that whole warps skip; inside, it is branch free
        int oc = fold1 ? 0 : 1;
        if (fold0 + fold1 != 1) oc = showdown_outcome();
        payoffs_given(out, oc);
    }
    __device__ __forceinline__ void payoffs_given(float *out, int outcome) const {
        const int w0 = outcome != 1, w1 = outcome != 0;
        const int pot = min(chips0, chips1);
        float p0 = 0.f;
        if (w0 != w1) p0 = w0 ? (float)pot : -(float)pot;
        out[0] = p0 * 0.5f; out[1] = -p0 * 0.5f;
    }
    // envs/limitholdem.py:40-71
    template <class T> __device__ __forceinline__ void encode_obs(int seat, bool reset_view, T *row) const {
        row[seat ? card[1] : card[0]] = (T)1; row[seat ? card[3] : card[2]] = (T)1;   // no dynamic index: card[] stays in registers
        const int np_ = n_public();
#pragma unroll
        for (int k = 0; k < 5; k++) if (k < np_) row[card[4 + k]] = (T)1;
        // the state object returned by reset() still references last episode's list (Q-LH1)
        const uint32_t shown = reset_view ? rn_shown : rn;
#pragma unroll
        for (int k = 0; k < 4; k++) row[52 + 5 * k + ((shown >> (3 * k)) & 7u)] = (T)1;
    }
};

// ========================================================================================
// No-limit Hold'em, 2 players, 100 chips each, 5 actions (games/nolimitholdem/{game,round,player}.py over
// games/limitholdem/{dealer,player,judger,utils}.py; envs/nolimitholdem.py).  4 game words:
//  w0: cards 0..4 (6 bits each): hole p0a, p1a, p0b, p1b (deal order game.py:73-74), flop0
//  w1: flop1, flop2, turn, river (6 bits each) | status0 [24:26) status1 [26:28) (0 alive 1 folded 2 all-in) |
//      pointer [28] | dealer_id + 1 [29:31) (0 = not drawn yet: the first init_game draws it and it is kept)
//  w2: in_chips0 [0:7) in_chips1 [7:14) remained0 [14:21) remained1 [21:28) round_counter [28:31)
//  w3: raised0 [0:7) raised1 [7:14) not_raise_num [14:16) not_playing_num [16:18) n_public [18:21)
// ========================================================================================
struct NoLimit {
    static constexpr int kGameId = 6, P = 2, A = 5, OBS = 54, GAME_WORDS = 4, MASK_WORDS = 1;
    static constexpr bool kUsesChain = false;
    static constexpr int kMaxResetDraws = 53;
    static constexpr int kRolloutMinEpw = 32;    // measured: fewer envs per warp only adds idle lanes (kernels.cuh)
    static constexpr bool kHasApply = false;
    static constexpr bool kWarpDeal = false;
    static constexpr bool kChanceAwareState = false;
    static constexpr int kSharedBytes = 0;
    static __device__ __forceinline__ void fill_shared(uint8_t *, int, int) {}
    __device__ __forceinline__ void bind_shared(const uint8_t *) {}
    enum { kFoldA = 0, kCheckCall = 1, kHalfPot = 2, kPot = 3, kAllIn = 4 };          // round.py:8-18
    int card[9];
    int in0, in1, rem0, rem1, st0, st1, pointer, dealer, rc, raised0, raised1, nrn, npn, npub;

    __device__ __forceinline__ void load(const uint32_t *st, size_t n, size_t i) {
        const uint32_t w0 = st[i], w1 = st[n + i], w2 = st[2 * n + i], w3 = st[3 * n + i];
#pragma unroll
        for (int k = 0; k < 5; k++) card[k] = bf_get(w0, 6 * k, 6);
#pragma unroll
        for (int k = 0; k < 4; k++) card[5 + k] = bf_get(w1, 6 * k, 6);
        st0 = bf_get(w1, 24, 2); st1 = bf_get(w1, 26, 2); pointer = bf_get(w1, 28, 1); dealer = (int)bf_get(w1, 29, 2) - 1;
        in0 = bf_get(w2, 0, 7); in1 = bf_get(w2, 7, 7); rem0 = bf_get(w2, 14, 7); rem1 = bf_get(w2, 21, 7); rc = bf_get(w2, 28, 3);
        raised0 = bf_get(w3, 0, 7); raised1 = bf_get(w3, 7, 7); nrn = bf_get(w3, 14, 2); npn = bf_get(w3, 16, 2); npub = bf_get(w3, 18, 3);
    }
    __device__ __forceinline__ void store(uint32_t *st, size_t n, size_t i) const {
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) w0 |= (uint32_t)card[k] << (6 * k);
#pragma unroll
        for (int k = 0; k < 4; k++) w1 |= (uint32_t)card[5 + k] << (6 * k);
        w1 |= (st0 << 24) | (st1 << 26) | (pointer << 28) | ((uint32_t)(dealer + 1) << 29);
        st[i] = w0; st[n + i] = w1;
        st[2 * n + i] = in0 | (in1 << 7) | (rem0 << 14) | (rem1 << 21) | (rc << 28);
        st[3 * n + i] = raised0 | (raised1 << 7) | (nrn << 14) | (npn << 16) | (npub << 18);
    }
    __device__ __forceinline__ void bet(int p, int chips) {                 // nolimitholdem/player.py:16-19
        const int r = p ? rem1 : rem0, q = min(chips, r);
        if (p) { in1 += q; rem1 -= q; } else { in0 += q; rem0 -= q; }
    }
    // game.py:50-111
    template <class Ch> __device__ __forceinline__ void reset(Ch &ch) {
        if (dealer < 0) dealer = (int)ch.below(2u);                        // drawn by the first init_game only (game.py:62-63)
        int j[9];
        if constexpr (Ch::kKind == 0) {
            const uint32_t x1 = ch.below(52u * 51u * 50u), x2 = ch.below(49u * 48u * 47u), x3 = ch.below(46u * 45u * 44u);
            j[0] = (int)(x1 / 2550u); j[1] = (int)((x1 / 50u) % 51u); j[2] = (int)(x1 % 50u);
            j[3] = (int)(x2 / 2256u); j[4] = (int)((x2 / 47u) % 48u); j[5] = (int)(x2 % 47u);
            j[6] = (int)(x3 / 1980u); j[7] = (int)((x3 / 44u) % 45u); j[8] = (int)(x3 % 44u);
        } else {
#pragma unroll
            for (int s = 0; s < 9; s++) j[s] = (int)ch.below((uint32_t)(52 - s));
            ch.skip_fy(42, 1);
        }
        fy_tail_cards<9>(52, j, card);
        in0 = in1 = 0; rem0 = rem1 = 100; st0 = st1 = 0; npub = 0;
        const int sb = (dealer + 1) & 1, bb = dealer;                      // (dealer + 2) % 2
        bet(bb, 2); bet(sb, 1);
        pointer = (bb + 1) & 1;
        nrn = 0; npn = 0; raised0 = in0; raised1 = in1; rc = 0;
    }
    __device__ __forceinline__ int pot() const { return in0 + in1; }       // dealer.pot as refreshed by get_state (game.py:186)
    // round.py:125-161
    __device__ __forceinline__ uint32_t legal_bits() const {
        const int mine = pointer ? raised1 : raised0, mx = max(raised0, raised1), rem = pointer ? rem1 : rem0;
        const int diff = mx - mine, pt = pot();
        uint32_t m = 0x1Fu;
        if (diff > 0 && diff >= rem) m &= ~((1u << kHalfPot) | (1u << kPot) | (1u << kAllIn));
        else {
            if (pt > rem) m &= ~(1u << kPot);
            if ((pt >> 1) > rem || (pt >> 1) + mine <= mx) m &= ~(1u << kHalfPot);
        }
        return m;
    }
    __device__ __forceinline__ void legal(uint32_t (&m)[1]) const { m[0] = legal_bits(); }
    __device__ __forceinline__ int player() const { return pointer; }
    __device__ __forceinline__ bool over() const { return ((st0 != 1) + (st1 != 1) == 1) || rc >= 4; }   // limitholdem/game.py:216-231
    // env.py:65-86, game.py:113-181, round.py:65-123
    template <class Ch> __device__ __forceinline__ void step(int id, Ch &, int &err) {
        const uint32_t m = legal_bits();
        if (id < 0 || id > 4 || !((m >> id) & 1u)) { id = kFoldA; err |= 4; }   // the reference has no fallback here
        const int p = pointer, pt = pot();
        const int mine = p ? raised1 : raised0, mx = max(raised0, raised1), rem = p ? rem1 : rem0;
        const bool fold = id == kFoldA, cc = id == kCheckCall;                  // branch free: lanes act differently
        const int add = id == kAllIn ? rem : (id == kPot ? pt : (id == kHalfPot ? (pt >> 1) : 0));
        const int amount = cc ? mx - mine : add;
        const int nr = cc ? mx : mine + add;
        nrn = cc ? nrn + 1 : (fold ? nrn : 1);
        if (p) raised1 = nr; else raised0 = nr;
        bet(p, amount);
        int stp = p ? st1 : st0;
        if (fold) stp = 1;
        if ((p ? rem1 : rem0) == 0 && stp != 1) stp = 2;
        if (p) st1 = stp; else st0 = stp;
        pointer = p ^ 1;
        if (stp == 2) { npn++; nrn--; }
        if (stp == 1) npn++;
        if ((pointer ? st1 : st0) == 1) pointer ^= 1;                      // skip the folded seat
        int by0 = st0 != 0, by1 = st1 != 0;                                // game.py:141-147
        if (by0 + by1 == 1) {
            const int last = by0 ? 1 : 0;
            if ((last ? raised1 : raised0) >= max(raised0, raised1)) { if (last) by1 = 1; else by0 = 1; }
        }
        if (nrn + npn >= 2) {                                              // round over: game.py:150-177
            const bool all = by0 + by1 == 2;
            pointer = (dealer + 1) & 1;
            if (!all && (pointer ? by1 : by0)) pointer ^= 1;
            if (rc == 0) { npub = 3; if (all) rc++; }
            if (rc == 1) { npub = 4; if (all) rc++; }
            if (rc == 2) { npub = 5; if (all) rc++; }
            rc++;
            nrn = 0; raised0 = raised1 = 0;
        }
    }
    // game.py:226-236 + limitholdem/judger.py for two players: the winner takes what the loser can match (chips, not blinds)
    __device__ __forceinline__ void payoffs(float *out) const {
        uint32_t bd[4] = {0, 0, 0, 0};                               // always scored, see Limit::payoffs
#pragma unroll
        for (int k = 4; k < 9; k++) holdem_add_card(bd, card[k]);
        uint32_t m0[4] = { bd[0], bd[1], bd[2], bd[3] }, m1[4] = { bd[0], bd[1], bd[2], bd[3] };
        holdem_add_card(m0, card[0]); holdem_add_card(m0, card[2]);
        holdem_add_card(m1, card[1]); holdem_add_card(m1, card[3]);
        const uint32_t s0 = holdem_strength_masks(m0), s1 = holdem_strength_masks(m1);
        const bool folded = (st0 == 1) != (st1 == 1);
        const int w0 = folded ? (st1 == 1) : (s0 >= s1), w1 = folded ? (st0 == 1) : (s1 >= s0);
        const int potm = min(in0, in1);
        float p0 = 0.f;
        if (w0 != w1) p0 = w0 ? (float)potm : -(float)potm;
        out[0] = p0; out[1] = -p0;
    }
    // envs/nolimitholdem.py:47-79: 52 card bits + my chips + max chips
    template <class T> __device__ __forceinline__ void encode_obs(int seat, bool, T *row) const {
        row[seat ? card[1] : card[0]] = (T)1; row[seat ? card[3] : card[2]] = (T)1;   // no dynamic index: card[] stays in registers
#pragma unroll
        for (int k = 0; k < 5; k++) if (k < npub) row[card[4 + k]] = (T)1;
        row[52] = (T)(seat ? in1 : in0);
        row[53] = (T)max(in0, in1);
    }
};

}  // namespace rlc
