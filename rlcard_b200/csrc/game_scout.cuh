// game_scout.cuh -- Scout, 4 players, 45 cards, 204 actions, one env per warp.
// Reference: rlcard/games/scout/{game,round,dealer,player,card,judger}.py, games/scout/utils/utils.py,
// rlcard/envs/scout.py.
//
// Hands and the table set are ordered lists of (top, bottom) cards with values 1..10, kept as two
// 16-nibble words each (slot k = bits [4k, 4k+4)), so that slicing, inserting and the group/run tests
// are a handful of 64-bit operations.  The state is replicated across the lanes (hands in the warp's
// shared memory, the rest in registers, see kernels_warp.cuh); for the
// legal mask (round.py:225-260) lane s owns the plays that start at hand position s -- one contiguous id
// range per start -- and the mask words are warp OR-reductions of those ranges.
//
// Game words (23): [0..7] hand tops p0..p3 (lo,hi)  [8..15] hand bottoms  [16,17] table tops
//   [18,19] table bottoms  [20] hl0..3 (5 bits each) | tl [20:25) | owner [25:28) (4 = none) |
//   consecutive_scouts [28:30) | current player [30:32)   [21] score0 | score1 << 16
//   [22] score2 | score3 << 16 (15 bits) | game_over << 31
#pragma once
#include "common.cuh"
#include "kernels_warp.cuh"

namespace rlc {

__device__ __forceinline__ uint64_t nib_mask(int len) { return len >= 16 ? ~0ull : ((1ull << (4 * len)) - 1ull); }
__device__ __forceinline__ uint64_t shr64(uint64_t x, int bits) { return bits >= 64 ? 0ull : (x >> bits); }
__device__ __forceinline__ uint64_t shl64(uint64_t x, int bits) { return bits >= 64 ? 0ull : (x << bits); }
__device__ __forceinline__ int nib(uint64_t x, int k) { return (int)((x >> (4 * k)) & 15ull); }

// LPE = lanes per env: 32 (a warp per env: replay modes, per-step kernels) or 16 (the fused throughput rollout hosts two
// envs per warp, one per half-warp: the scalar transition every lane replicates is then paid once per TWO envs, and the
// lane-parallel parts -- 16 starts, 16 hand / table slots -- are 16 wide anyway).  `lane` is always the lane within the
// env's group.  With LPE = 16 legal_words() and step() must be called by ALL 32 lanes together (k_wrollout_multi does):
// their collectives name the whole warp with a compile-time mask -- a per-lane group mask would make the compiler wrap
// every ballot / reduction in a MATCH.ANY + WARPSYNC loop -- and each half picks its own part of the result.  reset()
// may run in one half only: it has no collective, only barriers on the group's own mask.
template <int LPE>
struct ScoutT {
    static_assert(LPE == 32 || LPE == 16 || LPE == 8, "a warp, a half-warp or a quarter-warp per env");
    static constexpr int kPasses = LPE >= 16 ? 1 : 16 / LPE;      // the 16 hand / table slots (and starts) a lane covers
    static constexpr int kLanes = LPE;
    template <int L> using WithLanes = ScoutT<L>;
    static constexpr int kRolloutLanes = 8;       // lanes per env of the fused throughput rollout (RLC_WROLLOUT_LPE=32 / 16 force a warp / half-warp per env)
    uint32_t gm; int gsh;                         // LPE == 16: lane mask of this env's half-warp and its first lane
    __device__ __forceinline__ void bind_group(uint32_t mask, int shift) { gm = mask; gsh = shift; }
    __device__ __forceinline__ uint32_t ballot(bool pred) const {
        if constexpr (LPE == 32) return __ballot_sync(kFull, pred);
        else return (__ballot_sync(kFull, pred) >> gsh) & ((1u << LPE) - 1u);
    }
    __device__ __forceinline__ uint32_t red_or(uint32_t v) const {
        if constexpr (LPE == 32) return __reduce_or_sync(kFull, v);
        else if constexpr (LPE == 16) {
            const uint32_t lo = __reduce_or_sync(kFull, gsh ? 0u : v), hi = __reduce_or_sync(kFull, gsh ? v : 0u);
            return gsh ? hi : lo;
        } else {                                                       // butterfly inside the group's eight lanes
            v |= __shfl_xor_sync(kFull, v, 1); v |= __shfl_xor_sync(kFull, v, 2); v |= __shfl_xor_sync(kFull, v, 4);
            return v;
        }
    }
    __device__ __forceinline__ void wsync() const { __syncwarp(); }                          // all 32 lanes are here
    __device__ __forceinline__ void gsync() const { __syncwarp(LPE == 32 ? kFull : gm); }    // only this env's lanes may be
    static constexpr int kGameId = 5, P = 4, A = 204, OBS = 688, GAME_WORDS = 23, MASK_WORDS = 7;
    static constexpr bool kMaskBitpacked = false;
    static constexpr int kMinBlocks = 7;          // resident 128-thread blocks per SM the rollout kernel is compiled for (71 registers, no spills;
                                                  // measured 5 / 6 / 7 / 8 blocks: 1.175 / 1.139 / 1.050 / 1.043 ms, 8 spills)
    static constexpr int kScratchBytes = 48 + 64 + 32;   // reset: 45-card deck | the four hands (tops, bottoms) | legal-set accumulator (5 words)
    static constexpr bool kRowFlushFull = true;   // 2752-byte rows: batched compile-time flush (1.235 -> 1.212 ms)
    static constexpr bool kMaskBulk = false;      // 204-byte mask rows are not whole 16-byte units
    uint64_t *hands;       // shared memory of the warp: [0..3] hand tops p0..p3, [4..7] hand bottoms.  Every lane writes the
                           // same values (the state is replicated), so no barrier is needed around these accesses; keeping
                           // them out of the registers drops the 64-bit four-way select chains and 16 registers per thread
    uint64_t tt, tb;
    uint32_t hlw, sc01, sc23;     // hand lengths, one byte per seat | scores, 16 bits per seat: a run-time seat is a shift, not a select chain
    int tl, owner, consec, cur, over_;
    bool forced;           // current_player_forced_scout: recomputed by every legal() (round.py:246)

    __device__ __forceinline__ uint64_t top_of(int p) const { return hands[p]; }
    __device__ __forceinline__ uint64_t bot_of(int p) const { return hands[4 + p]; }
    __device__ __forceinline__ int hl_of(int p) const { return (int)((hlw >> (8 * p)) & 255u); }
    __device__ __forceinline__ void hl_set(int p, int v) { hlw = (hlw & ~(255u << (8 * p))) | ((uint32_t)v << (8 * p)); }
    __device__ __forceinline__ int score_of(int p) const { return (int)(((p & 2) ? sc23 : sc01) >> (16 * (p & 1)) & 0xffffu); }
    __device__ __forceinline__ void score_add(int p, int v) {
        const uint32_t d = (uint32_t)v << (16 * (p & 1));
        sc01 += (p & 2) ? 0u : d; sc23 += (p & 2) ? d : 0u;
    }

    uint32_t *acc_sm;      // LPE < 32: the lanes OR their id ranges into five shared words (two ATOMS per start) instead of 5 x 2 REDUX / 15 shuffles
    __device__ __forceinline__ void bind(const KParams &, uint8_t *scratch) {
        hands = reinterpret_cast<uint64_t *>(scratch + 48);
        acc_sm = reinterpret_cast<uint32_t *>(scratch + 48 + 64);
    }
    __device__ void load(const uint32_t *w, int lane) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
            hands[p] = (uint64_t)w[2 * p] | ((uint64_t)w[2 * p + 1] << 32);
            hands[4 + p] = (uint64_t)w[8 + 2 * p] | ((uint64_t)w[9 + 2 * p] << 32);
        }
        tt = (uint64_t)w[16] | ((uint64_t)w[17] << 32); tb = (uint64_t)w[18] | ((uint64_t)w[19] << 32);
        const uint32_t m = w[20], s0 = w[21], s1 = w[22];
#pragma unroll
        hlw = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) hlw |= ((m >> (5 * p)) & 31u) << (8 * p);
        tl = (m >> 20) & 31; owner = (m >> 25) & 7; consec = (m >> 28) & 3; cur = (m >> 30) & 3;
        sc01 = s0; sc23 = s1 & 0x7fffffffu; over_ = s1 >> 31;
        forced = false; lm_valid = false;
    }
    __device__ void store(uint32_t *w, int lane) const {
        if (lane != 0) return;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            w[2 * p] = (uint32_t)hands[p]; w[2 * p + 1] = (uint32_t)(hands[p] >> 32);
            w[8 + 2 * p] = (uint32_t)hands[4 + p]; w[9 + 2 * p] = (uint32_t)(hands[4 + p] >> 32);
        }
        w[16] = (uint32_t)tt; w[17] = (uint32_t)(tt >> 32); w[18] = (uint32_t)tb; w[19] = (uint32_t)(tb >> 32);
        w[20] = hl_of(0) | (hl_of(1) << 5) | (hl_of(2) << 10) | (hl_of(3) << 15) | (tl << 20) | (owner << 25) | (consec << 28) | ((uint32_t)cur << 30);
        w[21] = sc01;
        w[22] = (sc23 & 0x7fffffffu) | ((uint32_t)over_ << 31);
    }

    // utils/utils.py:17-67 + 144-184 on a nibble slice: valid?, (type 2 group / 1 run / 0 single, rank)
    static __device__ __forceinline__ bool segment(uint64_t tops, int s, int len, int &type, int &rank) {
        const uint64_t mk = nib_mask(len), seg = shr64(tops, 4 * s) & mk;
        const int first = (int)(seg & 15ull), last = (int)(shr64(seg, 4 * (len - 1)) & 15ull);
        const uint64_t rep = (uint64_t)first * (0x1111111111111111ull & mk), ramp = 0xFEDCBA9876543210ull & mk;
        const bool group = seg == rep, asc = seg == rep + ramp, desc = seg == rep - ramp;
        type = len == 1 ? 0 : (group ? 2 : 1);
        rank = (len > 1 && !group) ? max(first, last) : first;
        return len == 1 || group || asc || desc;
    }
    // round.py:225-260: 204-bit legal set of the current player, identical in all lanes.
    // Play ids are enumerated start-major (utils.py:186-200): start s owns the contiguous ids base(s) + len - 1,
    // base(s) = 16 s - s (s - 1) / 2.  From s the valid lengths are a prefix 1 .. 1 + max(#equal, #ascending, #descending
    // adjacent steps) (utils.py:17-67), cut at the hand's end; against a table set of tl cards the longer ones always
    // beat it and length tl does iff its (type, rank) is higher (round.py:262-295) -- so each start contributes ONE id
    // range.  Lane s works out its range from three adjacency ballots and the five mask words are OR-reductions of the
    // lanes' ranges; the scout ids 136 + 4 ins + {front, front flipped, back, back flipped} are a closed-form pattern.
    __device__ void legal_words(uint32_t (&m)[7], int lane) {
        const uint64_t T = top_of(cur);
        const int n = hl_of(cur);
        int ttype = 0, trank = 0;
        if (tl > 0) segment(tt, 0, tl, ttype, trank);
        // adjacency of hand positions (s, s + 1), s = 0 .. 14, as three 16-bit sets; a lane covers the positions lane + q LPE
        uint32_t eqm = 0, upm = 0, dnm = 0;
#pragma unroll
        for (int q = 0; q < kPasses; q++) {
            const int s = (lane & 15) + q * LPE;
            const int a = nib(T, s & 15), b = nib(T, (s + 1) & 15);
            const bool adj = lane < 16 && s < 15 && s + 1 < n;
            eqm |= ballot(adj && b == a) << (q * LPE);
            upm |= ballot(adj && b == a + 1) << (q * LPE);
            dnm |= ballot(adj && b == a - 1) << (q * LPE);
        }
        uint32_t acc[5] = {0u, 0u, 0u, 0u, 0u};
        if constexpr (LPE < 32) { if (lane < 5) acc_sm[lane] = 0u; wsync(); }
#pragma unroll
        for (int q = 0; q < kPasses; q++) {
            const int s = lane + q * LPE;
            int lo_id = 1, hi_id = 0;                                      // empty
            if (s < n && s < 16) {
                const int a = nib(T, s);
                const int req = __ffs((int)~(eqm >> s)) - 1, rup = __ffs((int)~(upm >> s)) - 1, rdn = __ffs((int)~(dnm >> s)) - 1;
                const int hi_len = min(max(req, max(rup, rdn)) + 1, n - s);
                int lo_len = 1;
                if (tl > 0) {                                              // a set of exactly tl cards: stronger than the table?
                    const bool group = tl - 1 <= req, asc = tl - 1 <= rup;
                    const int type = tl == 1 ? 0 : (group ? 2 : 1);
                    const int rank = (tl > 1 && !group) ? (asc ? a + tl - 1 : a) : a;      // run: max(first, last)
                    lo_len = (type > ttype || (type == ttype && rank > trank)) ? tl : tl + 1;   // validity of length tl: tl <= hi_len
                }
                const int base = 16 * s - ((s * (s - 1)) >> 1);
                lo_id = base + lo_len - 1; hi_id = base + hi_len - 1;
            }
            if constexpr (LPE < 32) {
                if (lo_id <= hi_id) {                                      // at most 16 ids: one or two words
                    const uint64_t bits = ((2ull << (hi_id - lo_id)) - 1ull) << (lo_id & 31);
                    atomicOr(acc_sm + (lo_id >> 5), (uint32_t)bits);
                    if ((uint32_t)(bits >> 32)) atomicOr(acc_sm + (lo_id >> 5) + 1, (uint32_t)(bits >> 32));
                }
            } else {
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const int x = max(lo_id - 32 * r, 0), y = min(hi_id - 32 * r, 31);
                    acc[r] |= x <= y ? ((2u << y) - 1u) & ~((1u << x) - 1u) : 0u;
                }
            }
        }
        if constexpr (LPE < 32) {
            wsync();
#pragma unroll
            for (int r = 0; r < 5; r++) m[r] = acc_sm[r];
        } else {
#pragma unroll
            for (int r = 0; r < 5; r++) m[r] = red_or(acc[r]);
        }
        forced = (m[0] | m[1] | m[2] | m[3] | m[4]) == 0;
        m[5] = m[6] = 0;
        if (tl > 0 && n < 16) {                                            // scout: insert position <= n; back variants need tl > 1
            const int nb = 4 * n + 4;
            const uint64_t S = (nb >= 64 ? ~0ull : (1ull << nb) - 1ull) & (tl > 1 ? ~0ull : 0x3333333333333333ull);
            m[4] |= (uint32_t)(S << 8); m[5] = (uint32_t)(S >> 24); m[6] = (uint32_t)(S >> 56);
        }
    }
    uint32_t lm[7];        // the last legal() set, identical in all lanes
    bool lm_valid;         // lm / forced already describe the current state (step() computed them)
    __device__ __forceinline__ int legal(uint32_t *smask, uint8_t *, int lane) {
        if constexpr (LPE == 32) { if (!lm_valid) legal_words(lm, lane); }
        else { if (__any_sync(kFull, !lm_valid)) legal_words(lm, lane); }   // warp-wide decision: the collectives need both halves
        lm_valid = true;
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 7; r++) smask[r] = lm[r];
        }
        return popc_words<7>(lm);
    }
    // k-th legal id in ascending order
    __device__ __forceinline__ int pick(const uint32_t *, const uint8_t *, int k, int) const { return kth_set_bit<7>(lm, k); }
    // the last legal() set in the order games/scout/round.py:225-260 builds it: plays of two or more cards (start-major,
    // utils/utils.py:83-95), then the single cards (:97-98), then the scouts by insert position (front, front-flip, back,
    // back-flip) -- i.e. ascending ids except that the singles (id = first id of their start) come after the longer plays
    __device__ void legal_order(const uint32_t *, int32_t *out, int stride, int lane) const {
        if (lane != 0) return;
        int n = 0;
        for (int pass = 0; pass < 2; pass++) {
            int single = 0;                                              // id of play-s-(s+1): 0, 16, 31, 45, ...
            for (int s = 0; s < 16; s++) {
                const int lo = pass ? single : single + 1, hi = pass ? single + 1 : single + 16 - s;
                for (int a = lo; a < hi; a++)
                    if ((lm[a >> 5] >> (a & 31)) & 1u) { if (n < stride) out[n] = a; n++; }
                single += 16 - s;
            }
        }
        for (int a = 136; a < 204; a++)
            if ((lm[a >> 5] >> (a & 31)) & 1u) { if (n < stride) out[n] = a; n++; }
        for (int k = n; k < stride; k++) out[k] = -1;
    }
    // games/scout/game.py:37-65, dealer.py:12-22, round.py:22-50: two shuffles (Q-SC1), round-robin deal
    template <class WCh> __device__ void reset(WCh &ch, uint8_t *deck, int lane) {
        if (lane == 0) {                                 // lane 0 owns the chance source: it shuffles alone, nothing is broadcast
            int n = 0;
            for (int top = 1; top <= 10; top++) for (int bot = top + 1; bot <= 10; bot++) deck[n++] = (uint8_t)((top << 4) | bot);
            for (int pass = 0; pass < 2; pass++)
                for (int i = 44; i >= 1; i--) {
                    const uint32_t j = ch.ch.below((uint32_t)i + 1u);
                    const uint8_t t = deck[i]; deck[i] = deck[j]; deck[j] = t;
                }
            deck[45] = (uint8_t)ch.ch.below(4u);         // first player (round.py:22-50), read by every lane below
        }
        gsync();
        uint64_t nt[4] = {0, 0, 0, 0}, nb[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 45; k++) {                   // deck.pop() = position 44-k -> player k%4, slot k/4
            const int c = deck[44 - k];
            nt[k & 3] |= (uint64_t)(c >> 4) << (4 * (k >> 2));
            nb[k & 3] |= (uint64_t)(c & 15) << (4 * (k >> 2));
        }
        cur = deck[45];
#pragma unroll
        for (int p = 0; p < 4; p++) { hands[p] = nt[p]; hands[4 + p] = nb[p]; }
        hlw = 0x0b0b0b0cu; sc01 = sc23 = 0;                                 // 12 cards to seat 0, 11 to the others
        gsync();
        tt = tb = 0; tl = 0; owner = 4; consec = 0; over_ = 0;
        forced = false; lm_valid = false;
    }
    __device__ __forceinline__ int player() const { return cur; }
    __device__ __forceinline__ bool over() const { return over_ != 0; }
    // env.py:65-86, envs/scout.py:53-80,130-133; round.py:63-153.  `forced` must be current (legal() ran).
    template <class WCh> __device__ void step(int id, WCh &, const uint32_t *smask, uint8_t *, int lane, int &err) {
        if (id < 0 || id >= A || !((smask[id >> 5] >> (id & 31)) & 1u)) {   // the reference raises; replay feeds legal ids
            err |= 4;
            id = kth_set_bit<7>(lm, 0);
            if (id < 0) return;
        }
        const int p = cur;
        const bool was_forced = forced;
        uint64_t T = top_of(p), B = bot_of(p);
        int n = hl_of(p);
        if (id < 136) {
            int s = 0, k = id;
            while (k >= 16 - s) { k -= 16 - s; s++; }
            const int e = s + k + 1, len = e - s;
            const uint64_t mk = nib_mask(len), lo = nib_mask(s);
            const uint64_t st = shr64(T, 4 * s) & mk, sb = shr64(B, 4 * s) & mk;
            T = (T & lo) | shl64(shr64(T, 4 * e), 4 * s);
            B = (B & lo) | shl64(shr64(B, 4 * e), 4 * s);
            n -= len;
            score_add(p, tl);
            tt = st; tb = sb; tl = len; owner = p; consec = 0;
        } else {
            const int k = id - 136, front = (k & 3) < 2, flip = k & 1;
            int ins = k >> 2;
            int ctop, cbot;
            if (front) { ctop = nib(tt, 0); cbot = nib(tb, 0); tt >>= 4; tb >>= 4; }
            else { ctop = nib(tt, tl - 1); cbot = nib(tb, tl - 1); tt &= nib_mask(tl - 1); tb &= nib_mask(tl - 1); }
            tl--;
            if (flip) { const int x = ctop; ctop = cbot; cbot = x; }
            if (ins > n) ins = n;                                           // list.insert clamps
            const uint64_t lo = nib_mask(ins);
            T = (T & lo) | ((uint64_t)ctop << (4 * ins)) | shl64(T & ~lo, 4);
            B = (B & lo) | ((uint64_t)cbot << (4 * ins)) | shl64(B & ~lo, 4);
            n++;
            if (was_forced && owner < 4) score_add(owner, 1);              // Q-SC2
            consec++;
            if (tl == 0) { owner = 4; consec = 0; }
            if (consec == 3 && owner < 4) over_ = 1;
        }
        wsync();                                                            // every lane has read the old hand
        hands[p] = T; hands[4 + p] = B; hl_set(p, n);
        wsync();
        cur = (p + 1) & 3;
        legal_words(lm, lane);                                              // next player cannot move -> round ends
        lm_valid = true;
        if ((lm[0] | lm[1] | lm[2] | lm[3] | lm[4] | lm[5] | lm[6]) == 0) over_ = 1;
    }
    __device__ __forceinline__ void payoffs(float *out) const {             // judger.py:15-30
#pragma unroll
        for (int p = 0; p < 4; p++) out[p] = (float)(score_of(p) - hl_of(p));
    }
    // envs/scout.py:171-235 (row pre-zeroed).  Scalars are float32(python float64 expression).
    template <class T> __device__ void encode_obs(int seat, bool, T *row, uint8_t *, int lane) const {
        if (lane < 16) {
#pragma unroll
            for (int q = 0; q < kPasses; q++) {
                const int s = lane + q * LPE;
                if (s < hl_of(seat)) {
                    row[s * 10 + nib(top_of(seat), s) - 1] = (T)1;
                    row[160 + s * 10 + nib(bot_of(seat), s) - 1] = (T)1;
                    row[640 + s] = (T)1;
                }
            }
        }
        if (LPE <= 16 || lane >= 16) {                                      // fewer than 32 lanes: the same lanes write the table slots
#pragma unroll
            for (int q = 0; q < kPasses; q++) {
                const int j = (lane & 15) + q * LPE;
                if (j < tl) {
                    row[320 + j * 10 + nib(tt, j) - 1] = (T)1;
                    row[480 + j * 10 + nib(tb, j) - 1] = (T)1;
                    row[656 + j] = (T)1;
                }
            }
        }
        if (lane < 16) {
#pragma unroll
            for (int q = 0; q < kPasses; q++) {
                const int i = lane + q * LPE;                               // scalar i of envs/scout.py:171-235
                float v = 0.f;
                if (i < 5) v = owner == i ? 1.f : 0.f;
                else if (i == 5) v = consec == 0 ? 0.f : (consec == 1 ? (float)(1.0 / 3.0) : (consec == 2 ? (float)(2.0 / 3.0) : 1.f));
                else if (i < 10) v = (float)hl_of(i - 6) * 0.0625f;
                else if (i == 10) v = (float)score_of(seat) * 0.0625f;
                else if (i == 11) v = (float)tl * 0.0625f;
                else if (i == 13) v = forced ? 1.f : 0.f;
                else if (i >= 14) {
                    const int r = tl == 0 ? 0 : (i == 14 ? nib(tt, 0) : nib(tt, tl - 1));
                    const float tenth[11] = { 0.f, (float)(1 / 10.0), (float)(2 / 10.0), (float)(3 / 10.0), (float)(4 / 10.0), (float)(5 / 10.0),
                                              (float)(6 / 10.0), (float)(7 / 10.0), (float)(8 / 10.0), (float)(9 / 10.0), 1.f };
                    v = tenth[r];
                }
                row[672 + i] = (T)v;
            }
        }
    }
};
using Scout = ScoutT<32>;

}  // namespace rlc
