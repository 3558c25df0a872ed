// game_blackjack.cuh -- Blackjack, 1 player, 1 deck (rlcard/games/blackjack/*, envs/blackjack.py).
// The ordered remaining deck is part of the state because deal_card() indexes it with
// np_random.choice(len(deck)) and pops that position (dealer.py:26-37).
#pragma once
#include "common.cuh"

namespace rlc {

// w0: p_sum [0:7) p_aces [7:10) d_first [10:16) d_vis_sum [16:23) d_vis_aces [23:26) winner [26:28)
//     (0 ongoing, 1 tie, 2 win, 3 lose; judger.py:25-52) | deck_len [28:32) is too small -> w1
// w1: deck_len;  w2..w14: deck bytes (replay modes)
// Throughput (Philox) mode: the shuffle is not drawn (deal_card picks a uniform index of the remaining
// deck anyway), so the deck stays in id order and is a 52-bit membership mask in w2,w3; the idx-th
// remaining card is the idx-th set bit.
struct Blackjack {
    static constexpr int kGameId = 0, P = 1, A = 2, OBS = 2, GAME_WORDS = 15, MASK_WORDS = 1;
    static constexpr bool kUsesChain = false;   // reset draws ride on the policy word (common.cuh chain())
    static constexpr int kSharedBytes = 0;
    static __device__ __forceinline__ void fill_shared(uint8_t *, int, int) {}
    __device__ __forceinline__ void bind_shared(const uint8_t *) {}
    static constexpr int kMaxResetDraws = 55;
    static constexpr int kRolloutMinEpw = 32;    // measured: fewer envs per warp only adds idle lanes (kernels.cuh)
    static constexpr bool kHasApply = false;
    static constexpr bool kWarpDeal = false;
    static constexpr bool kChanceAwareState = true;   // load_k/store_k<chance kind>
    int p_sum, p_aces, d_first, d_vis_sum, d_vis_aces, winner, deck_len;
    uint8_t deck[52];
    uint32_t mlo, mhi;

    template <int KIND> __device__ void load_k(const uint32_t *st, size_t n, size_t i) {
        const uint32_t w = st[i];
        p_sum = bf_get(w, 0, 7); p_aces = bf_get(w, 7, 3); d_first = bf_get(w, 10, 6); d_vis_sum = bf_get(w, 16, 7);
        d_vis_aces = bf_get(w, 23, 3); winner = bf_get(w, 26, 2); deck_len = (int)st[n + i];
        if constexpr (KIND == 0) { mlo = st[2 * n + i]; mhi = st[3 * n + i]; }
        else {
#pragma unroll
            for (int k = 0; k < 13; k++) {
                const uint32_t v = st[(size_t)(2 + k) * n + i];
                deck[4 * k] = v & 255u; deck[4 * k + 1] = (v >> 8) & 255u; deck[4 * k + 2] = (v >> 16) & 255u; deck[4 * k + 3] = v >> 24;
            }
        }
    }
    template <int KIND> __device__ void store_k(uint32_t *st, size_t n, size_t i) const {
        st[i] = p_sum | (p_aces << 7) | (d_first << 10) | (d_vis_sum << 16) | (d_vis_aces << 23) | (winner << 26);
        st[n + i] = (uint32_t)deck_len;
        if constexpr (KIND == 0) { st[2 * n + i] = mlo; st[3 * n + i] = mhi; }
        else {
#pragma unroll
            for (int k = 0; k < 13; k++)
                st[(size_t)(2 + k) * n + i] = deck[4 * k] | (deck[4 * k + 1] << 8) | (deck[4 * k + 2] << 16) | ((uint32_t)deck[4 * k + 3] << 24);
        }
    }
    // judger.py:54-73: A=11, T/J/Q/K=10; card id = 13*suit + rank (A=0, 2..9, T, J, Q, K)
    __device__ static __forceinline__ int card_score(int c) { const int r = c % 13; return r == 0 ? 11 : (r >= 9 ? 10 : r + 1); }
    // judger.py:66-72 `while score > 21 and count_a > 0: count_a -= 1; score -= 10` in closed form (no data-dependent loop under
    // divergence): the loop runs min(aces, ceil((sum - 21) / 10)) times
    __device__ static __forceinline__ int soft(int sum, int aces) {
        const int need = sum > 21 ? (sum - 12) / 10 : 0;
        return sum - 10 * min(aces, need);
    }
    template <class Ch> __device__ int deal(Ch &ch) {            // dealer.py:26-37
        const int idx = (int)ch.below((uint32_t)deck_len);
        deck_len--;
        if constexpr (Ch::kKind == 0) {                          // idx-th remaining card of the id-ordered deck
            const int nlo = __popc(mlo);
            const bool hi = idx >= nlo;
            const int b = nth_set_bit32(hi ? mhi : mlo, hi ? idx - nlo : idx);
            mlo &= hi ? 0xffffffffu : ~(1u << b); mhi &= hi ? ~(1u << b) : 0xffffffffu;
            return b + (hi ? 32 : 0);
        } else {
            const int c = deck[idx];
            for (int k = idx; k < deck_len; k++) deck[k] = deck[k + 1];
            return c;
        }
    }
    __device__ __forceinline__ int p_score() const { return soft(p_sum, p_aces); }
    __device__ __forceinline__ int d_score() const {
        return soft(d_vis_sum + card_score(d_first), d_vis_aces + (d_first % 13 == 0));
    }
    __device__ __forceinline__ void add_p(int c) { p_sum += card_score(c); p_aces += (c % 13 == 0); }
    __device__ __forceinline__ void add_d(int c) { d_vis_sum += card_score(c); d_vis_aces += (c % 13 == 0); }
    // game.py:22-54, dealer.py:6-24
    template <class Ch> __device__ void reset(Ch &ch) {
        if constexpr (Ch::kKind == 0) { mlo = 0xffffffffu; mhi = 0xfffffu; }   // shuffle not drawn (see header)
        else {
            for (int i = 0; i < 52; i++) deck[i] = (uint8_t)i;
            shuffle_tail_u8(ch, deck, 52, 52);
        }
        deck_len = 52;
        p_sum = p_aces = d_vis_sum = d_vis_aces = 0; winner = 0;
        add_p(deal(ch)); d_first = deal(ch); add_p(deal(ch)); add_d(deal(ch));
    }
    __device__ __forceinline__ void legal(uint32_t (&m)[1]) const { m[0] = 3u; }      // envs/blackjack.py:55
    __device__ __forceinline__ int player() const { return 0; }
    __device__ __forceinline__ bool over() const { return winner != 0; }             // game.py:192-205
    template <class Ch> __device__ void finish(Ch &ch) {                             // game.py:80-88,94-102
        while (d_score() < 17) add_d(deal(ch));
        const int ps = p_score(), ds = d_score();
        if (ps > 21) winner = 3;
        else if (ds > 21) winner = 2;
        else winner = ps > ds ? 2 : (ps < ds ? 3 : 1);
    }
    template <class Ch> __device__ void step(int id, Ch &ch, int &) {                // game.py:56-123
        if (id != 1) { add_p(deal(ch)); if (p_score() > 21) finish(ch); }
        else finish(ch);
    }
    __device__ __forceinline__ void payoffs(float *out) const {                      // envs/blackjack.py:62-78
        out[0] = winner == 2 ? 1.f : (winner == 1 ? 0.f : -1.f);
    }
    template <class T> __device__ __forceinline__ void encode_obs(int, bool, T *row) const {   // envs/blackjack.py:38-60
        row[0] = (T)p_score();
        row[1] = (T)(winner != 0 ? d_score() : soft(d_vis_sum, d_vis_aces));
    }
};

}  // namespace rlc
