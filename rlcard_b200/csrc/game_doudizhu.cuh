// game_doudizhu.cuh -- DouDizhu, 3 players, 27 472 actions, one env per warp.
// Reference: rlcard/games/doudizhu/{game,round,player,dealer,judger,utils}.py, rlcard/envs/doudizhu.py.
//
// Suits never matter: a hand is 15 rank counts (3456789TJQKA2BR) packed as nibbles in one 64-bit word.
// The legal set is the reference's, restated over the constant action table (SURVEY F-DDZ1/F-DDZ2):
//   lead   : every non-pass action whose rank counts are contained in the hand (== the judger's cached
//            playable_cards, judger.py:124-331)
//   follow : 'pass' + contained actions of the target's type with larger weight + bombs (larger bombs if
//            the target is a bomb) + rocket; only 'pass' against a rocket (utils.py:225-262)
// Containment of a table row in the hand is one SWAR test: (((hand | M) - row) & M) == M, M = 0x88..8
// (counts <= 4 < 8, so no borrow crosses a nibble).  The 32 lanes test 32 consecutive action ids and one
// __ballot_sync is one word of the bit-packed mask.  A word-level prefilter (nibble-wise minimum of the
// 32 rows of a word must be contained, and the word must overlap the candidate id ranges) skips most of
// the 859 words, so a step reads a few KB of the 220 KB table instead of all of it.
//
// Game words (20): [0..5] hands p0..p2 (lo,hi)  [6..11] played counts p0..p2
//   [12..16] last nine action ids of the trace, oldest first, 16 bits each (initialised to 'pass' = zero
//            rows; [16] upper half unused)   [17] last action of p0 | p1 << 16   [18] last action of p2 |
//            greater player's action << 16   [19] greater [0:2) (3 = none) | current [2:4) | winner [4:6) (3 = none)
#pragma once
#include "common.cuh"
#include "kernels_warp.cuh"

namespace rlc {

struct DdzTables {
    const uint64_t *rows;      // [27472] nibble-packed rank counts ('pass' = 0)
    const ulonglong2 *need;    // [864 + 32] per mask word j: .x = nibble-wise min over its 32 rows, .y = bit t set when an id of
                               //         the word has type t; entry 864 + b describes batch b of 32 mask words (1024 ids), b < 27
    const uint8_t *type;       // [27472]
    const uint8_t *weight;     // [27472]
    const uint32_t *tw_start;  // [39][17] first id of type t with weight >= w; [t][16] = end of type t
    const uint16_t *therm;     // [27472][27] the 54-byte obs block of every action (envs/doudizhu.py:153-167: 4 x 13 thermometer + two joker
                               //         bytes), built on the device from `rows` at upload: the twelve action blocks of a byte obs row are COPIED
};
constexpr int kDdzPass = 27471, kDdzRocket = 27470, kDdzBomb0 = 27457, kDdzTypeBomb = 35, kDdzTypeRocket = 36;
constexpr uint64_t kNibHi = 0x8888888888888888ull;

__device__ __forceinline__ bool ddz_contains(uint64_t hand, uint64_t row) { return (((hand | kNibHi) - row) & kNibHi) == kNibHi; }
// Type prefilter.  Whatever its kickers, every action of a type contains `len` consecutive ranks (within 3..A when
// len > 1) of at least `mult` cards each (doudizhu_table.py type_requirements, checked against all 27 472 rows by the
// CPU tests); the rocket is both jokers.  A hand that lacks that core holds no action of the type, so whole id ranges
// are skipped before any table row is fetched -- on a lead this leaves a handful of the ~48 mask words whose
// nibble-wise minimum alone cannot rule them out (four_two_* / trio_* words mix many ranks, their minimum is zero).
// Type ids: 0 solo 1 pair 2 trio 3 trio_solo 4 trio_pair | 5..12 solo_chain_5..12 | 13..20 pair_chain_3..10 |
// 21..25 trio_chain_2..6 | 26..29 trio_solo_chain_2..5 | 30..32 trio_pair_chain_2..4 | 33 four_two_solo
// 34 four_two_pair 35 bomb | 36 rocket | 37 pass.
__device__ __forceinline__ bool ddz_type_ok(uint64_t hand, int t) {
    if (t >= 36) return t == 36 && ((hand >> 52) & 15u) != 0u && ((hand >> 56) & 15u) != 0u;
    const int mult = t < 5 ? min(t + 1, 3) : (t < 13 ? 1 : (t < 21 ? 2 : (t < 33 ? 3 : 4)));
    const int len = t < 5 ? 1 : (t < 13 ? t : (t < 21 ? t - 10 : (t < 26 ? t - 19 : (t < 30 ? t - 24 : (t < 33 ? t - 28 : 1)))));
    uint64_t g = ((hand | kNibHi) - (uint64_t)mult * 0x1111111111111111ull) & kNibHi;     // bit 3 of nibble r: count_r >= mult
    if (len > 1) {
        g &= 0x0000888888888888ull;                                                        // chains live on ranks 3..A
        for (int i = 1; i < len; i++) g &= g >> 4;
    }
    return g != 0;
}
__device__ __forceinline__ int ddz_cards(uint64_t c) {           // number of cards = sum of the 15 nibbles
    const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    const uint32_t s = (lo & 0x0f0f0f0fu) + ((lo >> 4) & 0x0f0f0f0fu) + (hi & 0x0f0f0f0fu) + ((hi >> 4) & 0x0f0f0f0fu);   // four byte sums <= 16
    return (int)__dp4a(s, 0x01010101u, 0u);
}

struct Doudizhu {
    static constexpr int kGameId = 4, P = 3, A = 27472, OBS = 912, GAME_WORDS = 20, MASK_WORDS = 860;   // 27 472 ids = 859 words, rows padded to 860 (3440 B: whole 16-byte units, so a row can leave as one bulk copy)
    static constexpr bool kMaskBitpacked = true;
    static constexpr int kMinBlocks = 7;          // resident 128-thread blocks per SM the rollout kernel is compiled for
    static constexpr int kScratchBytes = 64 + 4 * 864 + 128 + 48;   // reset: 54-card deck | legal(): list of non-empty mask words | encode_obs: 16 count words | hands, played piles
    static constexpr int kObsScratch = 64 + 4 * 864;
    static constexpr bool kRowFlushFull = false;  // the batched flush spills at the 72-register cap (measured 2 % slower)
    static constexpr int kBulkDefault = 3;        // mask + obs rows as bulk copies: 1.350 -> 1.218 ms (profiles/r02_ddz_bulk.md)
    static constexpr bool kMaskBulk = true;       // the fused rollout may hand the 3440-byte mask row to the copy engine (kernels_warp.cuh)
    int n_legal, n_live; bool has_pass;                  // summary of the last legal() (warp-uniform)
    DdzTables tab;
    uint64_t *hand, *played;   // [3] each, in the warp's shared memory: every lane writes the same values (replicated state),
                               // dynamic seat index = one LDS instead of a 64-bit select chain, 12 registers fewer
    uint32_t tr[5];            // nine 16-bit action ids, entry k in tr[k/2] >> (16*(k&1)), k = 8 is the newest
    uint32_t last_by[3], greater_action;
    int greater, cur, winner;

    __device__ __forceinline__ uint64_t sel3(const uint64_t *a, int p) const { return a[p]; }
    __device__ __forceinline__ void put3(uint64_t *a, int p, uint64_t v) { a[p] = v; }
    __device__ __forceinline__ uint32_t trace_at(int k) const {
        const uint32_t w = k < 2 ? tr[0] : (k < 4 ? tr[1] : (k < 6 ? tr[2] : (k < 8 ? tr[3] : tr[4])));
        return (w >> (16 * (k & 1))) & 0xffffu;
    }
    __device__ __forceinline__ void bind(const KParams &p, uint8_t *scratch) {
        hand = reinterpret_cast<uint64_t *>(scratch + kObsScratch + 128); played = hand + 3;
        tab.rows = reinterpret_cast<const uint64_t *>(p.tab[0]); tab.need = reinterpret_cast<const ulonglong2 *>(p.tab[1]);
        tab.type = reinterpret_cast<const uint8_t *>(p.tab[2]); tab.weight = reinterpret_cast<const uint8_t *>(p.tab[3]);
        tab.tw_start = reinterpret_cast<const uint32_t *>(p.tab[4]); tab.therm = reinterpret_cast<const uint16_t *>(p.tab[5]);
        n_legal = 0; n_live = 0; has_pass = false;           // legal() relies on these describing the (zeroed) smask; writing the
                                                             // mask row as zero words + listed words on top measured 14 % slower
    }
    __device__ void load(const uint32_t *w, int) {
#pragma unroll
        for (int p = 0; p < 3; p++) {
            hand[p] = (uint64_t)w[2 * p] | ((uint64_t)w[2 * p + 1] << 32);
            played[p] = (uint64_t)w[6 + 2 * p] | ((uint64_t)w[7 + 2 * p] << 32);
        }
#pragma unroll
        for (int k = 0; k < 5; k++) tr[k] = w[12 + k];
        last_by[0] = w[17] & 0xffffu; last_by[1] = w[17] >> 16; last_by[2] = w[18] & 0xffffu; greater_action = w[18] >> 16;
        greater = w[19] & 3; cur = (w[19] >> 2) & 3; winner = (w[19] >> 4) & 3;
    }
    __device__ void store(uint32_t *w, int lane) const {
        if (lane != 0) return;
#pragma unroll
        for (int p = 0; p < 3; p++) {
            w[2 * p] = (uint32_t)hand[p]; w[2 * p + 1] = (uint32_t)(hand[p] >> 32);
            w[6 + 2 * p] = (uint32_t)played[p]; w[7 + 2 * p] = (uint32_t)(played[p] >> 32);
        }
#pragma unroll
        for (int k = 0; k < 5; k++) w[12 + k] = tr[k];
        w[17] = last_by[0] | (last_by[1] << 16); w[18] = last_by[2] | (greater_action << 16);
        w[19] = (uint32_t)greater | (cur << 2) | (winner << 4);
    }
    // game.py:23-51, round.py:25-39, dealer.py:12-76: rank-sorted deck, one shuffle, 17/17/17 + 3 to seat 0
    template <class WCh> __device__ void reset(WCh &ch, uint8_t *deck, int lane) {
        if (lane == 0) {                                   // lane 0 owns the chance source: it shuffles alone, nothing is broadcast
            for (int i = 0; i < 52; i++) deck[i] = (uint8_t)(i >> 2);
            deck[52] = 13; deck[53] = 14;
            for (int i = 53; i >= 1; i--) {
                const uint32_t j = ch.ch.below((uint32_t)i + 1u);
                const uint8_t t = deck[i]; deck[i] = deck[j]; deck[j] = t;
            }
        }
        __syncwarp();
        // rank counts of the three hands: lane l adds cards l and l + 32 into the owner's nibble (17 / 17 / 17, the last three
        // to the landlord); a count never exceeds 4, so the six 32-bit halves are plain warp sums
        uint32_t inc[3][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}};
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int k = lane + 32 * q;
            if (k < 54) {
                const int r = deck[k], owner = k < 51 ? k / 17 : 0;
                const uint32_t bit = 1u << (4 * (r & 7));
#pragma unroll
                for (int o = 0; o < 3; o++) { inc[o][0] += (owner == o && r < 8) ? bit : 0u; inc[o][1] += (owner == o && r >= 8) ? bit : 0u; }
            }
        }
        uint64_t nh[3];
#pragma unroll
        for (int o = 0; o < 3; o++)
            nh[o] = (uint64_t)__reduce_add_sync(kFull, inc[o][0]) | ((uint64_t)__reduce_add_sync(kFull, inc[o][1]) << 32);
#pragma unroll
        for (int q = 0; q < 3; q++) { hand[q] = nh[q]; played[q] = 0; }
        __syncwarp();
        const uint32_t pp = (uint32_t)kDdzPass | ((uint32_t)kDdzPass << 16);
        tr[0] = tr[1] = tr[2] = tr[3] = tr[4] = pp;
        last_by[0] = last_by[1] = last_by[2] = kDdzPass; greater_action = kDdzPass;
        greater = 3; cur = 0; winner = 3;
    }
    __device__ __forceinline__ int player() const { return cur; }
    __device__ __forceinline__ bool over() const { return winner != 3; }      // game.py:155-163

    // player.py:60-76 / game.py:110-128: bit-packed legal set of the current player into smask[859].
    // Also leaves, for pick(): the number of legal actions and a list of the non-empty mask words with the
    // number of legal ids before each (word index | prefix << 10), so the policy never rescans the mask.
    __device__ int legal(uint32_t *smask, uint8_t *scratch, int lane) {
        uint32_t *slist = reinterpret_cast<uint32_t *>(scratch + 64);
        // smask is all zero outside the words the previous legal() listed (kernels zero it once, bind() starts the list
        // empty): clear those instead of all 859 words
        for (int idx = lane; idx < n_live; idx += 32) smask[slist[idx] & 1023u] = 0;
        if (has_pass && lane == 0) smask[kDdzPass >> 5] = 0;
        __syncwarp();
        n_legal = 0; n_live = 0; has_pass = false;
        if (winner != 3) return 0;                                             // terminal: actions = []
        const uint64_t H = sel3(hand, cur);
        const bool lead = greater == 3 || greater == cur;
        // candidate id ranges [lo1,hi1) u [lo2,hi2) (+ rocket, handled as part of range 2 when adjacent)
        int lo1 = 0, hi1 = kDdzPass, lo2 = 0, hi2 = 0, tt_follow = kDdzTypeRocket;
        if (!lead) {
            const int tt = __ldg(tab.type + (greater_action)), tw = __ldg(tab.weight + (greater_action));
            tt_follow = tt;
            if (tt == kDdzTypeRocket) { lo1 = hi1 = 0; }
            else if (tt == kDdzTypeBomb) { lo1 = (int)__ldg(tab.tw_start + (tt * 17 + tw + 1)); hi1 = kDdzRocket + 1; }   // larger bombs + rocket
            else { lo1 = (int)__ldg(tab.tw_start + (tt * 17 + tw + 1)); hi1 = (int)__ldg(tab.tw_start + (tt * 17 + 16)); lo2 = kDdzBomb0; hi2 = kDdzRocket + 1; }
        }
        // which types the hand can play at all (bit t); a follow only asks about the target's type, bombs and the rocket
        uint64_t F;
        {
            const uint64_t quad = ddz_type_ok(H, 35) ? 1ull : 0ull, rocket = ddz_type_ok(H, kDdzTypeRocket) ? 1ull : 0ull;
            if (lead) {
                F = (uint64_t)__ballot_sync(kFull, ddz_type_ok(H, lane)) | ((ddz_type_ok(H, 32) ? 1ull : 0ull) << 32) |
                    (quad * 7ull << 33) | (rocket << 36);
            } else {
                F = (quad << 35) | (rocket << 36);
                if (tt_follow < kDdzTypeBomb) F |= (ddz_type_ok(H, tt_follow) ? 1ull : 0ull) << tt_follow;
            }
        }
        // level 0: one lane per batch of 32 mask words (1024 ids): in range, a playable type in it, nibble-wise minimum contained?
        bool blive = false;
        if (lane < (MASK_WORDS + 31) / 32) {
            const int a0 = 1024 * lane, a1 = a0 + 1024;
            const ulonglong2 nd = __ldg(tab.need + (864 + lane));
            blive = ((a0 < hi1 && a1 > lo1) || (a0 < hi2 && a1 > lo2)) && (nd.y & F) != 0 && ddz_contains(H, nd.x);
        }
        uint32_t lb = __ballot_sync(kFull, blive);
        int cnt = 0, nl = 0;
        while (lb) {
            const int wb = __ffs(lb) - 1; lb &= lb - 1;
            const int j = wb * 32 + lane;                                      // level 1: which words can hold a legal id
            bool live = false;
            if (j < MASK_WORDS) {
                const int a0 = 32 * j, a1 = a0 + 32;
                const bool overlap = (a0 < hi1 && a1 > lo1) || (a0 < hi2 && a1 > lo2);
                const ulonglong2 nd = __ldg(tab.need + (j));
                live = overlap && (nd.y & F) != 0 && ddz_contains(H, nd.x);
            }
            uint32_t lw = __ballot_sync(kFull, live);
            while (lw) {                                                       // level 2: expand the live words
                const int b = __ffs(lw) - 1; lw &= lw - 1;
                const int jj = wb * 32 + b, id = 32 * jj + lane;
                const bool cand = (id >= lo1 && id < hi1) || (id >= lo2 && id < hi2);
                const bool ok = cand && ddz_contains(H, __ldg(tab.rows + (id)));
                const uint32_t word = __ballot_sync(kFull, ok);
                if (word) {
                    if (lane == 0) { smask[jj] = word; slist[nl] = (uint32_t)jj | ((uint32_t)cnt << 10); }
                    nl++; cnt += __popc(word);
                }
            }
        }
        if (!lead) {
            if (lane == 0) smask[kDdzPass >> 5] |= 1u << (kDdzPass & 31);
            has_pass = true;
        }
        __syncwarp();
        n_live = nl; n_legal = cnt + (has_pass ? 1 : 0);
        return n_legal;
    }
    // ascending ids (the reference's lead order is the iteration order of a Python set of strings: undefined)
    __device__ void legal_order(const uint32_t *smask, int32_t *out, int stride, int lane) const {
        if (lane != 0) return;
        int n = 0;
        for (int w = 0; w < MASK_WORDS; w++) {
            uint32_t v = smask[w];
            while (v) { const int b = __ffs(v) - 1; v &= v - 1; if (n < stride) out[n] = 32 * w + b; n++; }
        }
        for (int k = n; k < stride; k++) out[k] = -1;
    }
    // k-th legal id in ascending order (0 <= k < n_legal of the last legal()); 'pass' is the largest id
    __device__ int pick(const uint32_t *smask, const uint8_t *scratch, int k, int lane) const {
        if (has_pass && k == n_legal - 1) return kDdzPass;
        const uint32_t *slist = reinterpret_cast<const uint32_t *>(scratch + 64);
        int c = 0;                                                            // entries whose prefix <= k (prefixes ascend)
        for (int base = 0; base < n_live; base += 32) {
            const int idx = base + lane;
            const uint32_t le = __ballot_sync(kFull, idx < n_live && (int)(slist[idx] >> 10) <= k);
            c += __popc(le);
            if (le != kFull) break;
        }
        const uint32_t ee = slist[c - 1];
        const int jj = (int)(ee & 1023u);
        return 32 * jj + nth_set_bit32(smask[jj], k - (int)(ee >> 10));
    }
    // env.py:65-86, game.py:53-81, round.py:52-79, player.py:78-108, judger.py:335-348
    template <class WCh> __device__ void step(int id, WCh &, const uint32_t *smask, uint8_t *scratch, int lane, int &err) {
        if (id < 0 || id >= A || !((smask[id >> 5] >> (id & 31)) & 1u)) {     // replay feeds legal ids only
            err |= 4;
            if (n_legal == 0) return;
            id = pick(smask, scratch, 0, lane);
        }
        const int p = cur;
        if (id != kDdzPass) {
            const uint64_t c = __ldg(tab.rows + (id));
            const uint64_t nhand = sel3(hand, p) - c, nplayed = sel3(played, p) + c;
            __syncwarp();                                                     // every lane has read the old piles
            put3(hand, p, nhand);
            put3(played, p, nplayed);
            __syncwarp();
            greater = p; greater_action = (uint32_t)id;
        }
        tr[0] = (tr[0] >> 16) | (tr[1] << 16); tr[1] = (tr[1] >> 16) | (tr[2] << 16);
        tr[2] = (tr[2] >> 16) | (tr[3] << 16); tr[3] = (tr[3] >> 16) | (tr[4] << 16);
        tr[4] = (uint32_t)id;                                                 // entry 8 = newest
        last_by[0] = p == 0 ? (uint32_t)id : last_by[0]; last_by[1] = p == 1 ? (uint32_t)id : last_by[1];
        last_by[2] = p == 2 ? (uint32_t)id : last_by[2];
        if (sel3(hand, p) == 0) winner = p;
        cur = p == 2 ? 0 : p + 1;
    }
    __device__ __forceinline__ void payoffs(float *out) const {              // judger.py:351-359 (landlord = seat 0)
        out[0] = winner == 0 ? 1.f : 0.f; out[1] = out[2] = winner == 0 ? 0.f : 1.f;
    }
    // envs/doudizhu.py:153-167: 54-d block of rank counts c; lane writes elements lane and lane+32
    template <class T> __device__ __forceinline__ void put54(T *dst, uint64_t c, int lane) const {
        const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);             // ranks 0..7 | ranks 8..14
        const int sh = 4 * (lane >> 2), k = lane & 3;                          // element lane: rank lane/4, copy lane%4
        dst[lane] = (T)(((lo >> sh) & 15u) > (uint32_t)k ? 1 : 0);
        if (lane < 22) {                                                       // element lane + 32: ranks 8..12, then B, R
            const uint32_t v = lane < 20 ? (uint32_t)(((hi >> sh) & 15u) > (uint32_t)k) : (uint32_t)(((hi >> (lane == 20 ? 20 : 24)) & 15u) != 0u);
            dst[lane + 32] = (T)v;
        }
    }
    __device__ __forceinline__ uint64_t action_counts(uint32_t id) const { return __ldg(tab.rows + (id)); }
    template <class T> __device__ __forceinline__ void one_hot(T *dst, int n, int size, int lane) const {   // :169-173 (Q-DDZ1)
        if (lane == 0) dst[n >= 1 ? n - 1 : size - 1] = (T)1;
    }
    // envs/doudizhu.py:26-91 (row pre-zeroed; 790 used for the landlord, 901 for peasants, stride 912)
    template <class T> __device__ void encode_obs(int seat, bool, T *row, uint8_t *scratch, int lane) const {
        const int up = seat == 2 ? 0 : seat + 1, down = seat == 0 ? 2 : seat - 1, mate = 3 - seat;
        if constexpr (sizeof(T) == 1) {
            // byte rows.  The twelve ACTION blocks (last action, the nine-action trace, and for peasants the landlord's and the
            // team-mate's last action) are 54-byte rows of the per-action table: lane l < 27 copies 16-bit unit l of each (ids are
            // warp-uniform and the trace index is a compile-time constant here, so there is no select chain and no arithmetic
            // per block).  The four COUNT blocks (own hand, the others' hands, two played piles) are expanded from their rank
            // counts: lane (r, parity) writes the 4-byte thermometer of rank r (r = 13: the two joker bytes) of two of them.
            uint8_t *rowb = reinterpret_cast<uint8_t *>(row);
            if (lane < 27) {
                const uint32_t newest = trace_at(8), prev = trace_at(7);
                const uint16_t *th = tab.therm + lane;
                uint16_t *d16 = reinterpret_cast<uint16_t *>(rowb) + lane;
                d16[27 * 2] = __ldg(th + 27u * (newest != (uint32_t)kDdzPass ? newest : prev));
#pragma unroll
                for (int k = 0; k < 9; k++) d16[27 * (3 + k)] = __ldg(th + 27u * trace_at(k));
                if (seat != 0) {
                    d16[27 * 14] = __ldg(th + 27u * last_by[0]);
                    d16[27 * 15] = __ldg(th + 27u * (mate == 1 ? last_by[1] : last_by[2]));
                }
            }
            uint64_t *cw = reinterpret_cast<uint64_t *>(scratch + kObsScratch);
            if (lane < 4) {                                                     // count blocks 0, 1, 12, 13 -> cw[0..3]
                uint64_t v = sel3(hand, seat);
                if (lane == 1) v = sel3(hand, up) + sel3(hand, down);
                else if (lane == 2) v = seat == 0 ? played[2] : played[0];
                else if (lane == 3) v = seat == 0 ? played[1] : sel3(played, mate);
                cw[lane] = v;
            }
            __syncwarp();
            const int r = lane & 15;
            if (r < 14) {
                const int par = lane >> 4;                                      // parity 0: blocks 0, 12 ; parity 1: blocks 1, 13
                const int sh = 4 * (r & 7);
                const uint32_t joker_bit = r == 13 ? 0x100u : 0u;      // r = 13 writes the two joker bytes (B in the low byte, R above)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int b = q == 0 ? par : 12 + par;
                    uint8_t *dst = rowb + 54 * b + 4 * r;
                    const uint32_t w = reinterpret_cast<const uint32_t *>(cw)[2 * (2 * q + par) + (r >> 3)], k = (w >> sh) & 15u;   // 0..4 copies of rank r
                    uint32_t v = __funnelshift_lc(0x01010101u, 0u, 8u * k);                             // k thermometer bytes
                    v |= (w >> 16) & joker_bit;                                                         // red joker: nibble 14 -> byte 1
                    *reinterpret_cast<uint16_t *>(dst) = (uint16_t)v;
                    if (r < 13) *reinterpret_cast<uint16_t *>(dst + 2) = (uint16_t)(v >> 16);
                }
            }
        } else {
            put54(row, sel3(hand, seat), lane);
            put54(row + 54, sel3(hand, up) + sel3(hand, down), lane);
            const uint32_t newest = trace_at(8), prev = trace_at(7);
            put54(row + 108, action_counts(newest != (uint32_t)kDdzPass ? newest : prev), lane);
#pragma unroll
            for (int k = 0; k < 9; k++) put54(row + 162 + 54 * k, action_counts(trace_at(k)), lane);
            if (seat == 0) {
                put54(row + 648, played[2], lane);
                put54(row + 702, played[1], lane);
            } else {
                put54(row + 648, played[0], lane);
                put54(row + 702, sel3(played, mate), lane);
                put54(row + 756, action_counts(last_by[0]), lane);
                put54(row + 810, action_counts(mate == 1 ? last_by[1] : last_by[2]), lane);
            }
        }
        if (seat == 0) {
            one_hot(row + 756, ddz_cards(hand[2]), 17, lane);
            one_hot(row + 773, ddz_cards(hand[1]), 17, lane);
        } else {
            one_hot(row + 864, ddz_cards(hand[0]), 20, lane);
            one_hot(row + 884, ddz_cards(sel3(hand, mate)), 17, lane);
        }
    }
};

}  // namespace rlc
