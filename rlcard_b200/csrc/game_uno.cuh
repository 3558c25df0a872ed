// game_uno.cuh -- UNO, 2 players, 108 cards, 61 actions, one env per thread.
// Reference: rlcard/games/uno/{game,round,dealer,card,utils}.py, rlcard/envs/uno.py.
//
// A card is its 6-bit code 15*colour + trait, which is the frozen UnoCard.str (card.py:22): for the
// eight wild cards the colour in the code is the ORIGINAL deck colour, so the code also identifies
// the physical wild card (Q-UNO1).  The mutable UnoCard.color only ever matters for the target, which
// is kept by value as (code, current colour).  Non-wild hand cards are interchangeable, so a hand is
// 52 two-bit counters plus two arrival-ordered lists of held wild / wild_draw_4 cards (Q-UNO2: the
// first held card of that trait is the one played, round.py:71-75).  Draw pile and played pile are
// ordered lists (replace_deck shuffles deck+played in list order, round.py:155-160).
//
// Game words (38): [0..26] 108 card bytes: draw pile at [0,dl), played pile stored downwards from 107
//   [27..30] seat 0 counters per colour (13 x 2 bits)   [31] seat 0 wild lists
//   [32..35] seat 1 counters                            [36] seat 1 wild lists
//   [37] dl[0:7) pl[7:14) target code[14:20) target colour[20:22) direction[22] current[23] winner+1[24:26) err8[26]
// wild list word: wild: count[0:3) entries[3:11) ; wild_draw_4: count[11:14) entries[14:22)
//
// BAG = true (throughput / Philox mode): the order of the draw pile is observable only through pop(), so
// "shuffle, then pop" is replaced by "pop a uniformly random remaining card": draw pile and played pile
// become multisets (2-bit counters per card code, one word per colour), shuffles draw nothing, a pop is
// r = below(pile size) resolved by cumulative counts in ascending code order.  Same distribution as the
// reference, ~20x fewer draws per episode and no per-thread card array.  Game words (19): [0..3] draw
// pile counters  [4..7] played pile counters  [8..11] seat 0 counters  [12] seat 0 wild lists  [13..16]
// seat 1 counters  [17] seat 1 wild lists  [18] meta (as word 37 above).
#pragma once
#include "common.cuh"

namespace rlc {

// Reverse Fisher-Yates over an identity deck of n cards of which only the last K positions are popped: original
// position of the card that ends at position n-1-s (same traceback as fy_tail_cards in game_poker.cuh).
template <int K>
__device__ __forceinline__ void fy_tail_positions(int n, const int (&j)[K], int (&c)[K]) {
#pragma unroll
    for (int s = 0; s < K; s++) {
        int pos = j[s];
#pragma unroll
        for (int r = s - 1; r >= 0; r--) pos = (pos == j[r]) ? n - 1 - r : pos;
        c[s] = pos;
    }
}
// physical position in the deck of utils.py:31-52 (27 cards per colour: 0, 1 1 .. 9 9, skip x2, reverse x2, draw_2 x2,
// wild, wild_draw_4) -> card code 15 * colour + trait
__device__ __forceinline__ int uno_code_at(int pos) {
    const int c = pos / 27, r = pos - 27 * c;
    const int t = r == 0 ? 0 : (r <= 18 ? (r + 1) >> 1 : (r <= 24 ? 10 + ((r - 19) >> 1) : r - 12));
    return 15 * c + t;
}

template <bool BAG>
struct UnoT {
    static constexpr int kGameId = 3, P = 2, A = 61, OBS = 240, GAME_WORDS = BAG ? 19 : 66, MASK_WORDS = 2;
    static constexpr bool kHasLegalOrder = !BAG;
    static constexpr bool kUsesChain = false;   // reset draws ride on the policy word (common.cuh chain())
    static constexpr bool kChanceAwareState = false;
    static constexpr int kSharedBytes = 0;
    static __device__ __forceinline__ void fill_shared(uint8_t *, int, int) {}
    __device__ __forceinline__ void bind_shared(const uint8_t *) {}
    static constexpr int kMaxResetDraws = 256;
    uint8_t cards[BAG ? 1 : 108];
    // replay modes also keep the hands as ARRIVAL-ORDERED card lists (seat 0 upwards from ho[0], seat 1 downwards from
    // ho[107]; the two hands never hold more than 108 cards): the reference's state['legal_actions'] is ordered by first
    // occurrence in the hand list (round.py:96-135), which only matters to callers that index it (RandomAgent)
    uint8_t ho[BAG ? 1 : 108];
    int hl0, hl1;
    uint32_t dk[4], pk[4];                         // BAG: draw / played pile counters (13 x 2 bits + wild, wd4 per colour)
    uint32_t hc[2][4], hw[2];
    int dl, pl, tcode, tcolor, dir, cur, winner;   // dir: 0 = +1, 1 = -1 ; winner: -1 none

    static constexpr int kBase = BAG ? 8 : 27;     // first hand word
    __device__ void load(const uint32_t *st, size_t n, size_t i) {
        if constexpr (BAG) {
#pragma unroll
            for (int c = 0; c < 4; c++) { dk[c] = st[(size_t)c * n + i]; pk[c] = st[(size_t)(4 + c) * n + i]; }
        } else {
#pragma unroll
            for (int k = 0; k < 27; k++) {
                const uint32_t v = st[(size_t)k * n + i];
                cards[4 * k] = v & 255u; cards[4 * k + 1] = (v >> 8) & 255u; cards[4 * k + 2] = (v >> 16) & 255u; cards[4 * k + 3] = v >> 24;
            }
        }
#pragma unroll
        for (int p = 0; p < 2; p++) {
#pragma unroll
            for (int c = 0; c < 4; c++) hc[p][c] = st[(size_t)(kBase + 5 * p + c) * n + i];
            hw[p] = st[(size_t)(kBase + 4 + 5 * p) * n + i];
        }
        const uint32_t m = st[(size_t)(kBase + 10) * n + i];
        dl = bf_get(m, 0, 7); pl = bf_get(m, 7, 7); tcode = bf_get(m, 14, 6); tcolor = bf_get(m, 20, 2);
        dir = bf_get(m, 22, 1); cur = bf_get(m, 23, 1); winner = (int)bf_get(m, 24, 2) - 1;
        if constexpr (!BAG) {
#pragma unroll
            for (int k = 0; k < 27; k++) {
                const uint32_t v = st[(size_t)(38 + k) * n + i];
                ho[4 * k] = v & 255u; ho[4 * k + 1] = (v >> 8) & 255u; ho[4 * k + 2] = (v >> 16) & 255u; ho[4 * k + 3] = v >> 24;
            }
            const uint32_t hl = st[(size_t)65 * n + i];
            hl0 = (int)(hl & 127u); hl1 = (int)((hl >> 7) & 127u);
        }
    }
    __device__ void store(uint32_t *st, size_t n, size_t i) const {
        if constexpr (BAG) {
#pragma unroll
            for (int c = 0; c < 4; c++) { st[(size_t)c * n + i] = dk[c]; st[(size_t)(4 + c) * n + i] = pk[c]; }
        } else {
#pragma unroll
            for (int k = 0; k < 27; k++)
                st[(size_t)k * n + i] = cards[4 * k] | (cards[4 * k + 1] << 8) | (cards[4 * k + 2] << 16) | ((uint32_t)cards[4 * k + 3] << 24);
        }
#pragma unroll
        for (int p = 0; p < 2; p++) {
#pragma unroll
            for (int c = 0; c < 4; c++) st[(size_t)(kBase + 5 * p + c) * n + i] = hc[p][c];
            st[(size_t)(kBase + 4 + 5 * p) * n + i] = hw[p];
        }
        st[(size_t)(kBase + 10) * n + i] = dl | (pl << 7) | (tcode << 14) | (tcolor << 20) | (dir << 22) | (cur << 23) | ((winner + 1) << 24);
        if constexpr (!BAG) {
#pragma unroll
            for (int k = 0; k < 27; k++)
                st[(size_t)(38 + k) * n + i] = ho[4 * k] | (ho[4 * k + 1] << 8) | (ho[4 * k + 2] << 16) | ((uint32_t)ho[4 * k + 3] << 24);
            st[(size_t)65 * n + i] = (uint32_t)hl0 | ((uint32_t)hl1 << 7);
        }
    }

    // ---- list primitives
    template <class Ch> __device__ __forceinline__ int pop_deck(Ch &ch, int &err) {   // deck.pop(); Q-UNO4: empty -> flag
        if (dl <= 0) { err |= 8; return -1; }
        if constexpr (BAG) {                                             // uniformly random remaining card
            int r = (int)ch.below((uint32_t)dl);
            dl--;
            // counters in unary (v = 1 -> 01, v = 2 -> 11): the r-th remaining card in ascending code order
            // is the r-th set bit of the four 30-bit words
            const uint32_t u0 = dk[0] | ((dk[0] >> 1) & 0x15555555u), u1 = dk[1] | ((dk[1] >> 1) & 0x15555555u);
            const uint32_t u2 = dk[2] | ((dk[2] >> 1) & 0x15555555u), u3 = dk[3] | ((dk[3] >> 1) & 0x15555555u);
            const int n0 = __popc(u0), n1 = n0 + __popc(u1), n2 = n1 + __popc(u2);
            const int c = (r >= n0) + (r >= n1) + (r >= n2);
            const uint32_t u = c == 0 ? u0 : (c == 1 ? u1 : (c == 2 ? u2 : u3));
            r -= c == 0 ? 0 : (c == 1 ? n0 : (c == 2 ? n1 : n2));
            const int t = nth_set_bit32(u, r) >> 1;
            const uint32_t dec = 1u << (2 * t);
            dk[0] -= c == 0 ? dec : 0u; dk[1] -= c == 1 ? dec : 0u; dk[2] -= c == 2 ? dec : 0u; dk[3] -= c == 3 ? dec : 0u;
            return 15 * c + t;
        } else { (void)ch; return cards[--dl]; }
    }
    __device__ __forceinline__ void unpop_deck(int code) {               // dealer.py:33-37 deck.append(top)
        if constexpr (BAG) {
            const int c = code / 15, t = code - 15 * c;
            const uint32_t inc = 1u << (2 * t);
            dk[0] += c == 0 ? inc : 0u; dk[1] += c == 1 ? inc : 0u; dk[2] += c == 2 ? inc : 0u; dk[3] += c == 3 ? inc : 0u;
            dl++;
        } else cards[dl++] = (uint8_t)code;
    }
    __device__ __forceinline__ void push_played(int code) {
        if constexpr (BAG) {
            const int c = code / 15, t = code - 15 * c;
            const uint32_t inc = 1u << (2 * t);
            pk[0] += c == 0 ? inc : 0u; pk[1] += c == 1 ? inc : 0u; pk[2] += c == 2 ? inc : 0u; pk[3] += c == 3 ? inc : 0u;
        } else cards[107 - pl] = (uint8_t)code;
        pl++;
    }
    __device__ __forceinline__ void order_append(int p, int code) {      // hand.append(card)
        if constexpr (!BAG) {
            if (p == 0) { if (hl0 + hl1 < 108) ho[hl0++] = (uint8_t)code; }
            else { if (hl0 + hl1 < 108) ho[107 - hl1++] = (uint8_t)code; }
        }
    }
    __device__ void order_remove_first(int p, int code) {               // hand.pop(index of the first card with this str)
        if constexpr (!BAG) {
            const int n = p ? hl1 : hl0;
            int at = -1;
            for (int k = 0; k < n && at < 0; k++) if (ho[p ? 107 - k : k] == code) at = k;
            if (at < 0) return;
            for (int k = at; k + 1 < n; k++) { if (p) ho[107 - k] = ho[107 - (k + 1)]; else ho[k] = ho[k + 1]; }
            if (p) hl1--; else hl0--;
        }
    }
    __device__ __forceinline__ void to_hand(int p, int code) {          // branch free: lanes hold different cards
        order_append(p, code);
        const int c = code / 15, t = code - 15 * c;
        const bool wild = t >= 13;
        hc_add(p, c, wild ? 0u : 1u << (2 * min(t, 12)));
        const int base = t == 13 ? 0 : 11;
        const uint32_t w = hwp(p);
        const int cnt = (w >> base) & 7;
        const uint32_t w2 = ((w & ~(3u << (base + 3 + 2 * cnt))) | ((uint32_t)c << (base + 3 + 2 * cnt))) + (1u << base);
        set_hw(p, wild ? w2 : w);
    }
    // register-friendly accessors (no dynamically indexed arrays -> no local memory)
    __device__ __forceinline__ uint32_t hcp(int p, int c_static) const { return p ? hc[1][c_static] : hc[0][c_static]; }
    __device__ __forceinline__ uint32_t hwp(int p) const { return p ? hw[1] : hw[0]; }
    __device__ __forceinline__ void set_hw(int p, uint32_t v) { hw[0] = p ? hw[0] : v; hw[1] = p ? v : hw[1]; }
    __device__ __forceinline__ void hc_add(int p, int c, uint32_t delta) {
        const int wi = 4 * p + c;
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int k = 0; k < 4; k++) hc[q][k] += (wi == 4 * q + k) ? delta : 0u;
    }
    __device__ __forceinline__ int wild_count(int p, int t) const { return (hwp(p) >> (t == 13 ? 0 : 11)) & 7; }
    __device__ __forceinline__ int take_first_wild(int p, int t) {        // original colour of the first held wild of trait t
        const int base = t == 13 ? 0 : 11;
        const uint32_t field = (hwp(p) >> base) & 0x7ffu;
        const int cnt = field & 7, first = (field >> 3) & 3;
        const uint32_t rest = ((field >> 5) << 3) | (uint32_t)(cnt - 1);  // drop entry 0, count - 1
        set_hw(p, (hwp(p) & ~(0x7ffu << base)) | ((rest & 0x7ffu) << base));
        return first;
    }
    __device__ __forceinline__ bool hand_empty(int p) const { return (hcp(p, 0) | hcp(p, 1) | hcp(p, 2) | hcp(p, 3)) == 0 && (hwp(p) & 0x3807u) == 0; }
    template <class Ch> __device__ void shuffle_deck(Ch &ch) {           // dealer.py:14-17 (BAG: nothing to draw)
        if constexpr (!BAG) shuffle_tail_u8(ch, cards, dl, dl);
    }
    template <class Ch> __device__ void replace_deck(Ch &ch) {            // round.py:155-160
        if constexpr (BAG) {
#pragma unroll
            for (int c = 0; c < 4; c++) { dk[c] += pk[c]; pk[c] = 0; }
            dl += pl; pl = 0;
        } else {
            uint8_t tmp[108];
            for (int k = 0; k < pl; k++) tmp[k] = cards[107 - k];
            for (int k = 0; k < pl; k++) cards[dl + k] = tmp[k];
            dl += pl; pl = 0;
            shuffle_deck(ch);
        }
    }
    template <class Ch> __device__ __forceinline__ void deal(int p, int num, Ch &ch, int &err) {   // dealer.py:19-26
        for (int k = 0; k < num; k++) { const int c = pop_deck(ch, err); if (c >= 0) to_hand(p, c); }
    }

    // games/uno/game.py:22-56, round.py:24-52, dealer.py:28-39, utils.py:31-52 (deck order)
    template <class Ch> __device__ void reset(Ch &ch) {
        int err = 0;
        if constexpr (BAG) {                                   // one 0, two of 1..9 / skip / reverse / draw_2, one wild, one wd4
#pragma unroll
            for (int c = 0; c < 4; c++) { dk[c] = 0x16AAAAA9u; pk[c] = 0; }
        } else {
            int n = 0;
            for (int c = 0; c < 4; c++) {
                for (int t = 0; t < 10; t++) { cards[n++] = (uint8_t)(15 * c + t); if (t) cards[n++] = (uint8_t)(15 * c + t); }
                for (int t = 10; t < 13; t++) { cards[n++] = (uint8_t)(15 * c + t); cards[n++] = (uint8_t)(15 * c + t); }
                cards[n++] = (uint8_t)(15 * c + 13); cards[n++] = (uint8_t)(15 * c + 14);
            }
        }
        dl = 108; pl = 0;
#pragma unroll
        for (int p = 0; p < 2; p++) { hc[p][0] = hc[p][1] = hc[p][2] = hc[p][3] = 0; hw[p] = 0; }
        hl0 = hl1 = 0;
        int top;
        if constexpr (BAG) {
            // throughput spec: the 15 cards of the opening deal come from a partial Fisher-Yates over the 108 physical
            // positions of the deck (15 draws and a traceback); what stays in the pile is a multiset anyway.  The fused
            // rollout deals with the whole warp instead (warp_deal below, same draws, same result).
            int j[15], pos[15];
#pragma unroll
            for (int q = 0; q < 15; q++) j[q] = (int)ch.below((uint32_t)(108 - q));
            fy_tail_positions<15>(108, j, pos);
            top = 0;
#pragma unroll
            for (int q = 0; q < 15; q++) {
                const int code = uno_code_at(pos[q]), c = code / 15, t = code - 15 * c;
                const uint32_t dec = 1u << (2 * t);
                dk[0] -= c == 0 ? dec : 0u; dk[1] -= c == 1 ? dec : 0u; dk[2] -= c == 2 ? dec : 0u; dk[3] -= c == 3 ? dec : 0u;
                if (q < 14) to_hand(q < 7 ? 0 : 1, code);
                else top = code;
            }
            dl = 93;
        } else {
            shuffle_deck(ch);
            deal(0, 7, ch, err); deal(1, 7, ch, err);
            top = pop_deck(ch, err);
        }
        open_round(ch, top, err);
    }
    // the rest of init_game once both hands and the first top card are dealt: dealer.py:28-39 (a wild_draw_4 goes back),
    // round.py:24-52 (colour of a wild top, opening effects)
    template <class Ch> __device__ void open_round(Ch &ch, int top, int &err) {
        dir = 0; cur = 0; winner = -1;
        while (top % 15 == 14) { unpop_deck(top); shuffle_deck(ch); top = pop_deck(ch, err); }
        tcode = top; tcolor = top / 15;
        if (top % 15 == 13) tcolor = (int)ch.below(4u);
        push_played(top);
        const int t = top % 15;                                            // round.py:38-52
        if (t == 10) cur = 1;
        else if (t == 11) { dir = 1; cur = 1; }
        else if (t == 12) deal(0, 2, ch, err);
    }
    // Opening deal of the fused rollout, by the whole warp: for every lane whose env starts an episode, lane q draws
    // swap partner q of the partial Fisher-Yates from that env's Philox stream, traces its card back through the
    // earlier swaps (14 shuffles), and the hands / pile counters are summed with warp reductions -- ~200 warp
    // instructions per deal instead of ~1900 executed by one lane while the other 31 wait.  Same draws and result as
    // reset() above.
    static constexpr bool kWarpDeal = BAG;
    __device__ void warp_deal(ChancePhilox &ch, bool starts, int lane) {
        uint32_t need = __ballot_sync(0xffffffffu, starts);
        while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            const uint32_t env = __shfl_sync(0xffffffffu, ch.env, L), kk = __shfl_sync(0xffffffffu, ch.k, L);
            const uint32_t dom = __shfl_sync(0xffffffffu, ch.dom, L), d0 = __shfl_sync(0xffffffffu, ch.d, L);
            const uint32_t k0 = __shfl_sync(0xffffffffu, ch.k0, L), k1 = __shfl_sync(0xffffffffu, ch.k1, L);
            const uint32_t idx = d0 + (uint32_t)lane;
            uint32_t o0, o1, o2, o3;
            philox4x32_10(kk, idx >> 2, env, dom, k0, k1, o0, o1, o2, o3);
            const int j = (int)__umulhi(sel4(o0, o1, o2, o3, idx & 3u), (uint32_t)(108 - min(lane, 14)));
            int pos = j;
#pragma unroll
            for (int r = 13; r >= 0; r--) {
                const int jr = __shfl_sync(0xffffffffu, j, r);
                pos = (r < lane && pos == jr) ? 107 - r : pos;
            }
            const int code = uno_code_at(pos), c = code / 15, t = code - 15 * c;
            const bool in0 = lane < 7, in1 = lane >= 7 && lane < 14, dealt = lane < 15;
            const uint32_t unit = 1u << (2 * t);
            uint32_t ndk[4], nh0[4], nh1[4], nhw[2] = {0u, 0u};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ndk[q] = __reduce_add_sync(0xffffffffu, (dealt && c == q) ? unit : 0u);
                nh0[q] = __reduce_add_sync(0xffffffffu, (in0 && c == q && t < 13) ? unit : 0u);
                nh1[q] = __reduce_add_sync(0xffffffffu, (in1 && c == q && t < 13) ? unit : 0u);
            }
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int tr = 13; tr < 15; tr++) {                     // held wilds: count + original colours in deal order
                    const bool mine = (p ? in1 : in0) && t == tr;
                    const uint32_t m = __ballot_sync(0xffffffffu, mine);
                    const int my = __popc(m & ((1u << lane) - 1u));
                    const uint32_t field = (uint32_t)__popc(m) | __reduce_or_sync(0xffffffffu, mine ? (uint32_t)c << (3 + 2 * my) : 0u);
                    nhw[p] |= field << (tr == 13 ? 0 : 11);
                }
            const int top = __shfl_sync(0xffffffffu, code, 14);
            if (lane == L) {
                int err = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) { dk[q] = 0x16AAAAA9u - ndk[q]; pk[q] = 0; hc[0][q] = nh0[q]; hc[1][q] = nh1[q]; }
                hw[0] = nhw[0]; hw[1] = nhw[1];
                dl = 93; pl = 0;
                ch.d = d0 + 15u;
                open_round(ch, top, err);
                ch.err |= err;
            }
        }
    }
    __device__ __forceinline__ int player() const { return cur; }
    __device__ __forceinline__ bool over() const { return winner >= 0; }   // game.py:154-160
    // round.py:96-135 as a 61-bit set
    __device__ void legal(uint32_t (&m)[2]) const {
        uint64_t bits = 0;
        const bool twild = (tcode % 15) >= 13;
        const int ttrait = tcode % 15;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t w = hcp(cur, c);
            // bit t of `have` <- 2-bit counter t of w is non-zero: OR the counter halves, then squeeze the even bits
            uint32_t have = (w | (w >> 1)) & 0x01555555u;
            have = (have | (have >> 1)) & 0x33333333u; have = (have | (have >> 2)) & 0x0f0f0f0fu;
            have = (have | (have >> 4)) & 0x00ff00ffu; have = (have | (have >> 8)) & 0x1fffu;
            uint32_t ok = (c == tcolor) ? have : 0u;
            if (!twild) ok |= have & (1u << ttrait);
            bits |= (uint64_t)ok << (15 * c);
        }
        if (wild_count(cur, 13)) bits |= (1ull << 13) | (1ull << 28) | (1ull << 43) | (1ull << 58);
        if (!bits && wild_count(cur, 14)) bits = (1ull << 14) | (1ull << 29) | (1ull << 44) | (1ull << 59);
        if (!bits) bits = 1ull << 60;
        m[0] = (uint32_t)bits; m[1] = (uint32_t)(bits >> 32);
    }
    // the same set in the insertion order of envs/uno.py:47-50 over round.py:96-135: cards in hand-list order, the four
    // colours of 'wild' where the first wild is met, a colour card where its first copy is met; wild_draw_4 only if
    // nothing else is legal, 'draw' only if nothing at all
    __device__ int legal_order(const uint32_t (&m)[2], int32_t *out, int stride) const {
        const uint64_t bits = (uint64_t)m[0] | ((uint64_t)m[1] << 32);
        uint64_t seen = 0;
        int n = 0;
        auto emit = [&](int a) { if (!((seen >> a) & 1ull)) { seen |= 1ull << a; if (n < stride) out[n] = a; n++; } };
        if constexpr (!BAG) {
            const int hn = cur ? hl1 : hl0;
            for (int k = 0; k < hn; k++) {
                const int code = ho[cur ? 107 - k : k], t = code % 15;
                if (t == 13) { if ((bits >> 13) & 1ull) { emit(13); emit(28); emit(43); emit(58); } }
                else if (t < 13 && ((bits >> code) & 1ull)) emit(code);
            }
        }
        for (int a = 0; a < 61; a++) if ((bits >> a) & 1ull) emit(a);     // wild_draw_4 x4 / draw, or ascending (multiset hands)
        return n;
    }
    // env.py:65-86, envs/uno.py:39-45, game.py:58-81, round.py:54-94 (play), 162-192 (draw), 194-227 (effects).
    // The lanes of a warp take different actions, so the transition is one mostly branch-free pass over flags; the only
    // branches left are the ones that make Philox draws (pop a card, colour of an auto-played wild, penalty cards).
    // smallest envs-per-warp the fused rollout may use on batches too small to fill the schedulers (measured: UNO gains
    // 3-12 % at 16 because half as many warp-steps pay for some lane's draw or reset; 8 loses it again to idle lanes)
    static constexpr int kRolloutMinEpw = 16;
    static constexpr bool kHasApply = true;      // the fused rollout only takes legal ids: it calls apply() directly
    template <class Ch> __device__ void step(int id, Ch &ch, int &err) {
        uint32_t m[2];
        legal(m);
        if (id < 0 || id > 60 || !((m[id >> 5] >> (id & 31)) & 1u)) {   // reference: random legal from the GLOBAL rng
            err |= 4; id = m[0] ? __ffs(m[0]) - 1 : 32 + __ffs(m[1]) - 1;
        }
        apply(id, ch, err);
    }
    template <class Ch> __device__ void apply(int id, Ch &ch, int &err) {
        const bool draw = id == 60;
        int card = 0;
        if (draw) {                                                     // _perform_draw_action: one card off the pile
            if (dl == 0) replace_deck(ch);
            card = pop_deck(ch, err);
        }
        const bool dead = draw && card < 0;                             // Q-UNO4: nothing left to draw, the turn just passes
        const int src = draw ? max(card, 0) : id;
        const int c = src / 15, t = src - 15 * c;                       // play: chosen colour; draw: the card's own colour
        const bool wild = t >= 13, play = !draw;
        const bool match = draw && !dead && !wild && c == tcolor;       // drawn card is played at once
        const bool auto_wild = draw && !dead && wild;                   // drawn wild is played with a random colour
        const bool keep = draw && !dead && !wild && c != tcolor;        // drawn card joins the hand
        // hand: played non-wild leaves it, kept card enters it, played wild = the first held one of that trait (Q-UNO2)
        const uint32_t unit = 1u << (2 * min(t, 12));
        hc_add(cur, c, (play && !wild) ? 0u - unit : (keep ? unit : 0u));
        int code = src;
        {
            const int base = t == 13 ? 0 : 11;
            const uint32_t w = hwp(cur), field = (w >> base) & 0x7ffu;
            const int cnt = field & 7, first = (field >> 3) & 3;
            const uint32_t rest = ((field >> 5) << 3) | (uint32_t)(cnt - 1);
            const bool take = play && wild;
            set_hw(cur, take ? (w & ~(0x7ffu << base)) | ((rest & 0x7ffu) << base) : w);
            code = take ? 15 * first + t : code;
        }
        if constexpr (!BAG) {
            if (keep) order_append(cur, src);
            if (play) order_remove_first(cur, code);
        }
        if (play && hand_empty(cur)) winner = cur;
        int color = c;
        if (auto_wild) color = (int)ch.below(4u);
        const bool to_pile = play || match || auto_wild;
        {                                                               // push_played, predicated
            const int pc = code / 15, pt = code - 15 * pc;
            if constexpr (BAG) {
                const uint32_t inc = to_pile ? 1u << (2 * pt) : 0u;
                pk[0] += pc == 0 ? inc : 0u; pk[1] += pc == 1 ? inc : 0u; pk[2] += pc == 2 ? inc : 0u; pk[3] += pc == 3 ? inc : 0u;
                pl += to_pile ? 1 : 0;
            } else if (to_pile) push_played(code);
        }
        tcode = to_pile ? code : tcode; tcolor = to_pile ? color : tcolor;
        // effects of a non-number card that was played from the hand or matched on the draw (never of an auto-played wild)
        const bool effect = (play || match) && t >= 10;
        const bool penalty = effect && (t == 12 || t == 14);
        if (penalty) {                                                  // draw_2 / wild_draw_4: the victim draws and is skipped
            const int need = t == 12 ? 2 : 4;
            if (dl < need) replace_deck(ch);
            deal(cur ^ 1, need, ch, err);
        }
        dir ^= (effect && t == 11) ? 1 : 0;
        const bool stay = effect && (t == 10 || penalty);               // two players: skip and penalties return the turn
        cur ^= stay ? 0 : 1;
    }
    __device__ __forceinline__ void payoffs(float *out) const {          // game.py:108-118
        out[0] = winner == 0 ? 1.f : (winner == 1 ? -1.f : 0.f);
        out[1] = winner == 1 ? 1.f : (winner == 0 ? -1.f : 0.f);
    }
    // encode_obs() of the acting seat by lane PAIRS: with 16 envs per warp (kernels.cuh) lanes 16..31 carry no env, so lane
    // L + 16 fetches env L's hand words by shuffle and writes the planes of colours 2, 3 while lane L writes colours 0, 1
    // and the target -- half the one-hot byte stores per lane.  Called by all 32 lanes; `row` is env (lane & 15)'s row.
    static constexpr bool kPairEmit = BAG;
    template <class T> __device__ void encode_obs_pair(bool valid, int lane, T *row) const {
        const int src = lane & 15, seat = cur;
        const uint32_t w0 = __shfl_sync(0xffffffffu, hcp(seat, 0), src), w1 = __shfl_sync(0xffffffffu, hcp(seat, 1), src);
        const uint32_t w2 = __shfl_sync(0xffffffffu, hcp(seat, 2), src), w3 = __shfl_sync(0xffffffffu, hcp(seat, 3), src);
        const uint32_t hwv = __shfl_sync(0xffffffffu, hwp(seat), src);
        const int tc = __shfl_sync(0xffffffffu, tcode, src);
        if (!__shfl_sync(0xffffffffu, (int)valid, src)) return;
        const bool upper = lane >= 16;
        const uint32_t wa = upper ? w2 : w0, wb = upper ? w3 : w1;
        T *ra = row + (upper ? 30 : 0), *rb = ra + 15;
#pragma unroll
        for (int t = 0; t < 13; t++) {
            ra[60 * ((wa >> (2 * t)) & 3) + t] = (T)1;                     // count 0 keeps plane 0 set
            rb[60 * ((wb >> (2 * t)) & 3) + t] = (T)1;
        }
        const int kw = (hwv & 7u) ? 60 : 0, k4 = ((hwv >> 11) & 7u) ? 60 : 0;
        ra[kw + 13] = (T)1; ra[k4 + 14] = (T)1; rb[kw + 13] = (T)1; rb[k4 + 14] = (T)1;
        if (!upper) row[180 + tc] = (T)1;
    }
    // envs/uno.py:24-33, utils.py:69-127 (row pre-zeroed): planes 0..2 hand copies, plane 3 target
    template <class T> __device__ void encode_obs(int seat, bool, T *row) const {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t w = hcp(seat, c);
#pragma unroll
            for (int t = 0; t < 13; t++) {
                const int k = (w >> (2 * t)) & 3;
                row[60 * k + 15 * c + t] = (T)1;                       // k == 0 keeps plane 0 set
            }
            const int kw = wild_count(seat, 13) ? 60 : 0, k4 = wild_count(seat, 14) ? 60 : 0;
            row[kw + 15 * c + 13] = (T)1;
            row[k4 + 15 * c + 14] = (T)1;
        }
        row[180 + tcode] = (T)1;                                       // target.str: original colour (Q-UNO1)
    }
};

using Uno = UnoT<false>;      // replay modes: ordered piles
using UnoBag = UnoT<true>;    // throughput mode

}  // namespace rlc
