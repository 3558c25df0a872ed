#!/usr/bin/env python
"""Harvest known-answer vectors for the judgers from the LIVE reference (build container only).

Sources:
  * hold'em 7-card evaluator: every ``compare_hands`` call the reference's own unit tests make
    (tests/utils/test_holdem_utils.py, tests/games/test_limitholdem_game.py) is recorded by running those
    tests with a recording wrapper around the function -- inputs and the (asserted) outputs -- plus random
    deals scored by the reference's ``compare_hands``;
  * DouDizhu judger: the in / not-in vectors of tests/games/test_doudizhu_judger.py (commented out there; all
    are checked here against the live engine before they are stored) with the full playable set of each hand
    from ``DoudizhuJudger.playable_cards_from_hand``, random hands, and random follow situations scored by
    ``get_gt_cards``;
  * Leduc judger: ``LeducholdemJudger.judge_game`` over every rank / chips / fold combination (includes the
    three cases of tests/games/test_leducholdem_game.py:62-91);
  * UNO encoders: ``encode_hand`` / ``encode_target`` on the vectors of tests/games/test_uno_game.py:72-102
    and on random hands.

Output: tests/golden/kats.npz.   Usage: python tests/golden/make_kats.py
"""
import ast
import os
import random
import re
import sys
import unittest

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402

RANKS = 'A23456789TJQK'
SUITS = {'S': 0, 'H': 1, 'D': 2, 'C': 3, 'B': 2}     # 'B' is the test file's stand-in for the fourth suit


def card_id(s):
    return 13 * SUITS[s[0]] + RANKS.index(s[1])


def harvest_holdem(rlcard):
    import importlib
    sys.path.insert(0, REF)
    rec = []
    mods = []
    for name in ('tests.utils.test_holdem_utils', 'tests.games.test_limitholdem_game'):
        m = importlib.import_module(name)
        if hasattr(m, 'compare_hands'):
            orig = m.compare_hands

            def wrapper(hands, _orig=orig):
                out = _orig(hands)
                rec.append(([None if h is None else list(h) for h in hands], list(out)))
                return out
            m.compare_hands = wrapper
            mods.append(m)
    suite = unittest.TestSuite()
    for m in mods:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(m))
    res = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, 'w')).run(suite)
    assert res.wasSuccessful(), res.failures + res.errors
    return rec


def main():
    rlcard = import_reference()
    from rlcard.games.limitholdem.utils import compare_hands
    rng = random.Random(20261018)
    out = {}

    # ---------------------------------------------------------------- hold'em evaluator
    vecs = [(h, w, 0) for h, w in harvest_holdem(rlcard)]
    n_ref = len(vecs)
    deck = [s + r for s in 'SHDC' for r in RANKS]
    for _ in range(30000):
        P = rng.choice([2, 2, 2, 3, 4])
        cards = rng.sample(deck, 5 + 2 * P)
        hands = [cards[5 + 2 * p:7 + 2 * p] + cards[:5] for p in range(P)]
        for p in range(P):
            if rng.random() < 0.1:
                hands[p] = None
        if all(h is None for h in hands):
            continue
        for h in hands:
            if h is not None:
                rng.shuffle(h)
        vecs.append((hands, list(compare_hands([None if h is None else list(h) for h in hands])), 1))
    keep = []
    for hands, win, src in vecs:
        if len(hands) > 4 or any(h is not None and len(h) != 7 for h in hands):
            continue
        letters = set(c[0] for h in hands if h for c in h)
        if 'B' in letters and 'D' in letters:
            continue
        keep.append((hands, win, src))
    N = len(keep)
    hc = np.full((N, 4, 7), 255, np.uint8)
    hn = np.zeros(N, np.int32); hw = np.zeros((N, 4), np.uint8); hs = np.zeros(N, np.uint8); hd = np.zeros(N, np.uint8)
    for i, (hands, win, src) in enumerate(keep):
        hn[i] = len(hands); hs[i] = src
        for p, h in enumerate(hands):
            hw[i, p] = win[p]
            if h is not None:
                ids = [card_id(c) for c in h]
                hc[i, p] = ids
                if len(set(ids)) != 7:
                    hd[i] = 1                                   # a physical card twice: cannot occur in a game
    out.update(holdem_cards=hc, holdem_np=hn, holdem_winners=hw, holdem_src=hs, holdem_dup=hd)
    print('holdem: %d vectors (%d recorded from the reference tests, %d usable, %d with duplicate cards)' % (
        N, n_ref, int((hs == 0).sum()), int(hd.sum())))

    # ---------------------------------------------------------------- DouDizhu judger
    from rlcard.games.doudizhu.judger import DoudizhuJudger as Judger
    from rlcard.games.doudizhu.utils import ACTION_2_ID, get_gt_cards
    from rlcard.games.doudizhu.utils import cards2str  # noqa: F401
    DR = '3456789TJQKA2BR'
    src = open(os.path.join(REF, 'tests', 'games', 'test_doudizhu_judger.py')).read()
    asserts, hands = [], []
    # the vectors are commented out in the reference file: strip the comment markers, then read the
    # ``in_cards / not_in_cards / hand`` literals (single-line and multi-line forms) -- parsed, never executed
    text = '\n'.join(re.sub(r'^\s*# ?', '', l) for l in src.splitlines() if l.strip().startswith('#'))
    found = []
    for m in re.finditer(r'in_cards, not_in_cards, hand = (.+)', text):
        found.append((m.start(), ast.literal_eval(m.group(1).strip())))
    for m in re.finditer(r"in_cards\s*=\s*(\(.*?\))\s*\n\s*not_in_cards\s*=\s*(\(.*?\))\s*\n\s*hand\s*=\s*('[^']*')", text, re.S):
        found.append((m.start(), tuple(ast.literal_eval(m.group(k)) for k in (1, 2, 3))))
    for _, (inc, notin, hand) in sorted(found, key=lambda t: t[0]):
        inc = (inc,) if isinstance(inc, str) else inc
        notin = (notin,) if isinstance(notin, str) else notin
        if hand not in hands:
            hands.append(hand)
        asserts.append((hands.index(hand), inc, notin))
    n_test_hands = len(hands)
    full = ''.join(r * 4 for r in DR[:13]) + 'BR'
    for _ in range(400):
        k = rng.randint(1, 20)
        hands.append(''.join(sorted(rng.sample(full, k), key=DR.index)))
    hands.append(full)                                        # the whole deck: every action is playable (test :146-156)
    cases = []                                                # (counts[15], target id or -1, set of legal ids)
    for h in hands:
        playable = Judger.playable_cards_from_hand(h)
        cases.append(([h.count(r) for r in DR], -1, set(ACTION_2_ID[a] for a in playable)))
    trip = []
    n_in = n_out = 0
    for hi, inc, notin in asserts:
        playable = Judger.playable_cards_from_hand(hands[hi])
        for a in inc:
            assert a in playable, (hands[hi], a)
            trip.append((hi, ACTION_2_ID[a], 1)); n_in += 1
        for a in notin:
            assert a not in playable, (hands[hi], a)
            if a in ACTION_2_ID:
                trip.append((hi, ACTION_2_ID[a], 0)); n_out += 1

    class P:                                                    # the two attributes get_gt_cards reads
        pass

    class C:                                                    # jokers are Card(suit='BJ'/'RJ', rank='') (base.py, utils.py:136-151)
        def __init__(self, r):
            self.rank, self.suit = ('', {'B': 'BJ', 'R': 'RJ'}[r]) if r in 'BR' else (r, 'S')

    ids = sorted(ACTION_2_ID.values())
    id2a = {v: k for k, v in ACTION_2_ID.items()}
    for _ in range(1200):
        k = rng.randint(1, 20)
        h = ''.join(sorted(rng.sample(full, k), key=DR.index))
        tgt = rng.choice(ids[:-1]) if rng.random() < 0.8 else rng.choice(ids[-16:-1])   # bombs / rocket more often
        me, gp = P(), P()
        me.current_hand = [C(r) for r in h]
        gp.played_cards = id2a[tgt]
        try:
            gt = get_gt_cards(me, gp)
        except Exception:
            continue
        cases.append(([h.count(r) for r in DR], tgt, set(ACTION_2_ID[a] for a in gt)))
    M = len(cases)
    dh = np.zeros((M, 15), np.uint8); dt = np.zeros(M, np.int32); dl = np.zeros((M, 27472), np.uint8)
    for i, (cnt, tgt, legal) in enumerate(cases):
        dh[i] = cnt; dt[i] = tgt; dl[i, sorted(legal)] = 1
    out.update(ddz_hand=dh, ddz_target=dt, ddz_legal=np.packbits(dl, axis=1, bitorder='little'),
               ddz_assert=np.asarray(trip, np.int32), ddz_n_test_hands=np.int32(n_test_hands))
    print('doudizhu: %d legal-set cases (%d hands of the reference test file, %d in / %d not-in assertions verified)' % (
        M, n_test_hands, n_in, n_out))

    # ---------------------------------------------------------------- Leduc judger
    from rlcard.games.leducholdem.judger import LeducholdemJudger
    from rlcard.games.leducholdem.player import LeducholdemPlayer
    from rlcard.games.base import Card
    rows, pays = [], []
    rs = np.random.RandomState(0)
    for r0 in range(3):
        for r1 in range(3):
            for pub in range(3):
                if (r0 == r1 == pub):
                    continue                                       # three cards of one rank do not exist
                for c0, c1 in ((1, 2), (2, 1), (4, 4), (6, 2), (10, 10), (14, 14), (3, 7)):
                    for f0, f1 in ((0, 0), (1, 0), (0, 1)):
                        ps = [LeducholdemPlayer(0, rs), LeducholdemPlayer(1, rs)]
                        ps[0].hand, ps[1].hand = Card('S', 'JQK'[r0]), Card('H', 'JQK'[r1])
                        ps[0].in_chips, ps[1].in_chips = c0, c1
                        ps[0].status, ps[1].status = ('folded' if f0 else 'alive'), ('folded' if f1 else 'alive')
                        pay = LeducholdemJudger(rs).judge_game(ps, Card('S' if pub != r0 else 'H', 'JQK'[pub]))
                        rows.append((r0, r1, pub, c0, c1, f0, f1)); pays.append(pay)
    out.update(leduc_case=np.asarray(rows, np.int32), leduc_payoffs=np.asarray(pays, np.float64))
    print('leduc: %d judge_game cases' % len(rows))

    # ---------------------------------------------------------------- UNO encoders
    from rlcard.games.uno.utils import encode_hand, encode_target
    colors, traits = 'rgby', [str(i) for i in range(10)] + ['skip', 'reverse', 'draw_2', 'wild', 'wild_draw_4']
    uhands = [['y-1', 'r-8', 'b-9', 'y-reverse', 'r-skip'], ['y-4', 'y-4', 'r-skip', 'r-skip'], ['r-wild', 'g-wild_draw_4']]
    pool = []
    for c in colors:
        for t in traits:
            pool += [c + '-' + t] * (1 if t in ('0', 'wild', 'wild_draw_4') else 2)
    for _ in range(300):
        uhands.append(rng.sample(pool, rng.randint(0, 25)))
    uh = np.full((len(uhands), 32), 255, np.uint8); ut = np.zeros(len(uhands), np.uint8); uo = np.zeros((len(uhands), 240), np.uint8)
    for i, h in enumerate(uhands):
        codes = [15 * colors.index(s.split('-')[0]) + traits.index(s.split('-')[1]) for s in h]
        uh[i, :len(codes)] = codes
        tgt = rng.choice(pool)
        ut[i] = 15 * colors.index(tgt.split('-')[0]) + traits.index(tgt.split('-')[1])
        plane = np.zeros((4, 4, 15), dtype=int)
        encode_hand(plane[:3], h)
        encode_target(plane[3], tgt)
        uo[i] = plane.reshape(-1)
    out.update(uno_hand=uh, uno_target=ut, uno_obs=uo)
    print('uno: %d encoder cases' % len(uhands))

    np.savez_compressed(os.path.join(HERE, 'kats.npz'), **out)
    print('wrote', os.path.join(HERE, 'kats.npz'), os.path.getsize(os.path.join(HERE, 'kats.npz')), 'bytes')


if __name__ == '__main__':
    main()
