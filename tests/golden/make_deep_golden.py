#!/usr/bin/env python
"""Deep replay fixtures from the LIVE reference engine (build container only; needs /root/reference).

Same recording method as make_golden.py (the reference's ``Env`` played through its public API with the engine's
``np_random`` wrapped), but sized so that a mistake shared by the C oracle and the CUDA engines in a rare state has a
witness: >= 10 000 episodes for blackjack / leduc / limit / no-limit hold'em, >= 1 000 for UNO, >= 500 for DouDizhu
and Scout, plus stratified slots:

  uniform      uniform-random legal ids (poker: 10 % raw ids that may be illegal -> the env's fallback), real
               ``np.random.RandomState`` seeded like ``rlcard.make(seed=...)``, draws recorded -> the slot also pins
               the seed -> MT19937 -> draw chain
  stratified   class-uniform policy: the legal ids are grouped into classes (DouDizhu: the 38 action types; Scout:
               play length / scout side x flip x insert at front, end, middle; UNO: card trait; poker: rarely fold;
               blackjack: mostly hit) and a class is drawn first, so rare action types are played on lead and follow
  crafted      DouDizhu deals forced through a scripted shuffle (the 519-lead hand, rocket + bombs, trio_pair_chain_4,
               trio_solo_chain_5, long chains, bomb-heavy peasants), played with the stratified policy
  reshuffle    UNO episodes (stratified policy) in which UnoRound.replace_deck ran at least twice

To keep the committed files small, obs rows and legal sets are stored as 32-bit digests
(blake2b(digest_size=4) of the float32 obs row / of the ascending uint16 legal ids); the small fixtures of
make_golden.py keep full rows for diagnosis.

Output tests/golden/deep_<game>.npz:
  slot_seed[S] (-1 = scripted chance), slot_kind[S] (0 uniform 1 stratified 2 crafted 3 reshuffle), slot_episodes[S]
  tape u8 + tape_off[S+1], rec_off[S+1]; per record: rec_kind (0 reset 1 step 2 get_state(seat) 3 payoffs), rec_arg,
  rec_player, rec_done, rec_obs_dim, rec_obs_hash u32, rec_legal_hash u32, rec_nlegal u16; pay_rec / pay (payoff rows)

Usage:  python tests/golden/make_deep_golden.py [game ...]
"""
import collections
import hashlib
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import TapeRecorder, import_reference  # noqa: E402


def h32(b):
    return int.from_bytes(hashlib.blake2b(b, digest_size=4).digest(), 'little')


def obs_hash(obs):
    return h32(np.ascontiguousarray(obs, dtype='<f4').tobytes())


def legal_hash(ids):
    return h32(np.asarray(sorted(ids), dtype='<u2').tobytes())


class ScriptedRandom:
    """np.random.RandomState stand-in whose bounded draws come from a script (then from a seeded fallback).  The
    draw -> call mapping is numpy's legacy one: shuffle(list) = reverse Fisher-Yates, j_i = below(i+1) for i = n-1..1;
    randint(lo, hi) = lo + below(hi-lo); choice(n | list) = below(n)."""

    def __init__(self, seed):
        self.fb = random.Random(seed)
        self.script = collections.deque()
        self.tape = []

    def below(self, n):
        v = self.script.popleft() if self.script else self.fb.randrange(n)
        assert 0 <= v < n, (v, n)
        self.tape.append(v)
        return v

    def shuffle(self, x):
        for i in range(len(x) - 1, 0, -1):
            j = self.below(i + 1)
            x[i], x[j] = x[j], x[i]

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        return low + self.below(high - low)

    def choice(self, a):
        if isinstance(a, (int, np.integer)):
            return self.below(int(a))
        a = list(a)
        return a[self.below(len(a))]


def draws_for_permutation(post):
    """post[i] = index (in the pre-shuffle list) of the element that must end at position i -> the Fisher-Yates draws."""
    n = len(post)
    cur = list(range(n))
    where = list(range(n))             # where[e] = current position of element e
    out = []
    for i in range(n - 1, 0, -1):
        j = where[post[i]]
        assert j <= i
        out.append(j)
        a, b = cur[i], cur[j]
        cur[i], cur[j] = b, a
        where[a], where[b] = j, i
    assert cur == list(post)
    return out


# ------------------------------------------------------------------------------------------------ policies
def classes_for(game, env, state, ids, ctx):
    """legal ids -> {class key: [ids]} for the stratified policy."""
    groups = collections.defaultdict(list)
    if game == 'doudizhu':
        for a in ids:
            s = ctx['id2action'][a]
            groups['pass' if s == 'pass' else ctx['card_type'][s][0][0]].append(a)
    elif game == 'scout':
        hand_len = len(state['raw_obs']['hand'])
        for a in ids:
            p = ctx['actions'][a].split('-')
            if p[0] == 'play':
                groups[('play', min(int(p[2]) - int(p[1]), 4))].append(a)
            else:
                ins = int(p[2])
                where = 'front' if ins == 0 else ('end' if ins >= hand_len else 'mid')
                groups[('scout', p[1], p[3], where)].append(a)
    elif game == 'uno':
        for a in ids:
            s = ctx['actions'][a]
            groups['draw' if s == 'draw' else s.split('-', 1)[1]].append(a)
    else:
        for a in ids:
            groups[a].append(a)
    return groups


def pick(game, kind, env, state, pol, ctx, p_raw):
    ids = list(state['legal_actions'].keys())
    if kind == 0:
        if p_raw and pol.random() < p_raw:
            return pol.randrange(env.num_actions)
        return ids[pol.randrange(len(ids))]
    if game == 'blackjack':
        return 0 if (0 in ids and pol.random() < 0.7) else ids[pol.randrange(len(ids))]
    if game in ('leduc-holdem', 'limit-holdem'):
        keep = [a for a in ids if a != 2] or ids                     # 2 = fold
        return 2 if (2 in ids and pol.random() < 0.04) else keep[pol.randrange(len(keep))]
    if game == 'no-limit-holdem':
        keep = [a for a in ids if a != 0] or ids                     # 0 = FOLD
        return 0 if (0 in ids and pol.random() < 0.04) else keep[pol.randrange(len(keep))]
    groups = classes_for(game, env, state, ids, ctx)
    keys = sorted(groups, key=str)
    if game == 'scout' and len(keys) > 1:
        # grow the hands: scouting is preferred 2:1 so that 15/16-card hands and end inserts occur
        sc = [k for k in keys if k[0] == 'scout']
        if sc and pol.random() < 0.66:
            keys = sc
    if game == 'doudizhu' and len(keys) > 1 and 'pass' in keys and pol.random() < 0.8:
        keys = [k for k in keys if k != 'pass']                      # follow when able
    if game == 'doudizhu' and pol.random() < 0.5:
        # coverage seeking: the class played least so far (lead and follow counted apart)
        follow = 'pass' in groups
        k = min(keys, key=lambda c: (ctx['seen'][(c, follow)], str(c)))
    else:
        k = keys[pol.randrange(len(keys))]
    if game == 'doudizhu':
        ctx['seen'][(k, 'pass' in groups)] += 1
    g = groups[k]
    return g[pol.randrange(len(g))]


# ------------------------------------------------------------------------------------------------ crafted DouDizhu deals
RANKS = '3456789TJQKA2BR'
CRAFTED_DDZ = [
    # (landlord 20 cards | None, peasant-1 17 cards | None)
    ('333444555666789TJQKA', None),            # 519 legal leads (ADVICE: widest set)
    ('BR2222AAAAKKKKQQQQJJ', None),            # rocket + four bombs + four_two_*
    ('3456789TJQKA3456789T', None),            # 12-chain + pair chains
    ('333444555666778899TT', None),            # trio_pair_chain_4 exactly 20 cards
    ('33334444555566667777', None),            # five bombs, four_two_pair
    ('TTTJJJQQQKKKAAA34567', None),            # trio_solo_chain_5
    ('333444555666777888BR', None),            # trio chain 6 + rocket kickers
    ('33445566778899TTJJQQ', None),            # pair chain 10
    (None, 'BR2222AAAAKKKKQQQ'),               # peasant after the landlord beats anything
    (None, '3333444455556666T'),               # peasant with four bombs
    ('34567893456789BR2222', 'TTTTJJJJQQQQKKKKA'),
    ('34567345673456722BRA', 'AAAKKKQQQJJJTTT99'),
    ('3334445556789TJQKA2B', '66677788899TTJJQQ'),   # trio chains / trio_solo_chain_3 / trio_pair_chain_3 on lead and follow
    ('3456789TJQ3456789TJQ', '456789TJQKA456789'),   # long solo / pair chains, higher ones behind
    ('333356789TJQKA2BR789', '44445566TTJJQQKKA'),   # four_two_* lead and follow
    ('3334445556667777JKA2', '888999TTTJJJQQQKK'),   # trio_solo_chain_4/5 material on both sides
    ('3334445556667778KA2B', '888999TTTJJJQQQKA'),
    ('3456789TJQK3456789TJ', '3456789TJQKA45678'),   # solo_chain_11 / pair_chain_9
]


def ddz_script(rlcard, pol, landlord, peasant1):
    from rlcard.games.doudizhu.dealer import DoudizhuDealer
    from rlcard.games.doudizhu.utils import cards2str
    deck = DoudizhuDealer(np.random.RandomState(0)).deck
    rank_of = [cards2str([c]) for c in deck]
    free = list(range(54))
    pol.shuffle(free)

    def take(ranks):
        got = []
        for r in ranks:
            k = next(i for i in free if rank_of[i] == r)
            free.remove(k)
            got.append(k)
        return got
    ll = take(landlord) if landlord else None
    p1 = take(peasant1) if peasant1 else None
    if ll is None:
        ll = [free.pop() for _ in range(20)]
    if p1 is None:
        p1 = [free.pop() for _ in range(17)]
    p2 = free
    assert len(ll) == 20 and len(p1) == 17 and len(p2) == 17
    pol.shuffle(ll)
    post = ll[:17] + p1 + p2 + ll[17:]
    return draws_for_permutation(post)


# ------------------------------------------------------------------------------------------------ UNO reshuffle scan
def _uno_scan_chunk(args):
    """seeds whose single stratified-policy episode calls UnoRound.replace_deck at least `want` times (no recording)"""
    lo, hi, want = args
    rl = import_reference()
    from rlcard.games.uno.round import UnoRound
    from rlcard.games.uno.utils import ACTION_LIST
    n = {'n': 0}
    orig = UnoRound.replace_deck

    def counted(self):
        n['n'] += 1
        return orig(self)
    UnoRound.replace_deck = counted
    ctx = {'actions': ACTION_LIST}
    hits = []
    for seed in range(lo, hi):
        env = rl.make('uno', config={'seed': seed})
        pol = random.Random(seed * 7919 + 1)
        n['n'] = 0
        state, _ = env.reset()
        while not env.is_over():
            state, _ = env.step(pick('uno', 1, env, state, pol, ctx, 0.0))
        if n['n'] >= want:
            hits.append((seed, n['n']))
    UnoRound.replace_deck = orig
    return hits


def scan_uno_reshuffles(base, count, want=2, chunk=2000, limit=400000):
    import multiprocessing as mp
    found = []
    with mp.get_context('fork').Pool(os.cpu_count() or 1) as pool:
        lo = base
        while len(found) < count and lo < base + limit:
            jobs = [(lo + i * chunk, lo + (i + 1) * chunk, want) for i in range(os.cpu_count() or 1)]
            for hits in pool.map(_uno_scan_chunk, jobs):
                found.extend(hits)
            lo += chunk * len(jobs)
    found.sort()
    return found[:count], lo - base


# ------------------------------------------------------------------------------------------------ recording
PLAN = {
    # game: [(slot kind, slots, episodes per slot, p_raw)]
    'blackjack': [(0, 100, 100, 0.0), (1, 20, 50, 0.0)],
    'leduc-holdem': [(0, 100, 100, 0.10), (1, 20, 50, 0.0)],
    'limit-holdem': [(0, 100, 100, 0.10), (1, 20, 50, 0.0)],
    'no-limit-holdem': [(0, 100, 100, 0.0), (1, 20, 50, 0.0)],
    'uno': [(0, 100, 10, 0.0), (1, 25, 4, 0.0), (3, 12, 1, 0.0)],
    'doudizhu': [(0, 50, 8, 0.0), (1, 25, 4, 0.0), (2, len(CRAFTED_DDZ), 6, 0.0)],
    'scout': [(0, 50, 8, 0.0), (1, 25, 5, 0.0)],
}


def record_game(rlcard, game, out_dir):
    t_start = time.time()
    probe = rlcard.make(game, config={'seed': 0})
    P, A = probe.num_players, probe.num_actions
    ctx = {}
    if game == 'doudizhu':
        from rlcard.games.doudizhu.utils import CARD_TYPE, ID_2_ACTION
        ctx = {'id2action': ID_2_ACTION, 'card_type': CARD_TYPE[0], 'seen': collections.Counter()}
    elif game == 'scout':
        from rlcard.envs.scout import ACTION_LIST
        ctx = {'actions': ACTION_LIST}
    elif game == 'uno':
        from rlcard.games.uno.utils import ACTION_LIST
        ctx = {'actions': ACTION_LIST}
    reshuffles = {'n': 0}
    if game == 'uno':
        from rlcard.games.uno.round import UnoRound
        if not hasattr(UnoRound, '_orig_replace_deck'):
            UnoRound._orig_replace_deck = UnoRound.replace_deck

            def counted(self):
                reshuffles['n'] += 1
                return UnoRound._orig_replace_deck(self)
            UnoRound.replace_deck = counted

    R = dict(kind=[], arg=[], player=[], done=[], dim=[], oh=[], lh=[], nl=[])
    pay_rec, pay = [], []
    tapes, seeds, kinds, eps, rec_off = [], [], [], [], [0]
    stats = collections.Counter()

    def emit(kind, arg, env, state, payoffs=None):
        R['kind'].append(kind); R['arg'].append(arg)
        R['player'].append(env.get_player_id()); R['done'].append(int(env.is_over()))
        if state is not None:
            o = np.asarray(state['obs']).reshape(-1)
            ids = list(state['legal_actions'].keys())
            R['dim'].append(o.size); R['oh'].append(obs_hash(o)); R['lh'].append(legal_hash(ids)); R['nl'].append(len(ids))
        else:
            R['dim'].append(0); R['oh'].append(0); R['lh'].append(0); R['nl'].append(0)
        if payoffs is not None:
            pay_rec.append(len(R['kind']) - 1); pay.append(np.asarray(payoffs, np.float64))

    gi = list(PLAN).index(game)
    candidate = 0
    for kind, slots, episodes, p_raw in PLAN[game]:
        made = 0
        scanned = None
        if kind == 3:                                          # seeds found by a scan without recording
            scanned, tried = scan_uno_reshuffles(100000 * (gi + 1) + 10000 * kind, slots)
            print('    uno reshuffle scan: %d episodes tried, hits %s' % (tried, scanned))
            slots = len(scanned)
        while made < slots:
            seed = scanned[made][0] if scanned else 100000 * (gi + 1) + 10000 * kind + candidate
            candidate += 1
            mark = {k: len(v) for k, v in R.items()}
            mark_pay = len(pay_rec)
            reshuffles['n'] = 0
            env = rlcard.make(game, config={'seed': seed})
            scripted = kind in (2,)
            rec = ScriptedRandom(seed) if scripted else TapeRecorder(env.np_random)
            env.np_random = rec
            env.game.np_random = rec
            pol = random.Random(seed * 7919 + 1)
            for ep in range(episodes):
                if kind == 2:
                    ll, p1 = CRAFTED_DDZ[made]
                    rec.script.extend(ddz_script(rlcard, pol, ll, p1))
                state, pid = env.reset()
                if kind == 2 and ep == 0 and CRAFTED_DDZ[made][0]:
                    got = env.game.players[0].initial_hand if hasattr(env.game.players[0], 'initial_hand') else None
                    want = ''.join(sorted(CRAFTED_DDZ[made][0], key=RANKS.index))
                    assert got == want, (got, want)
                emit(0, 0, env, state)
                while not env.is_over():
                    a = pick(game, 0 if kind == 0 else 1, env, state, pol, ctx, p_raw)
                    if game == 'doudizhu':
                        s = ctx['id2action'][a] if a in state['legal_actions'] else None
                        if s:
                            follow = any(ctx['id2action'][x] == 'pass' for x in state['legal_actions'])
                            stats[('pass' if s == 'pass' else ctx['card_type'][s][0][0]) + ('/f' if follow else '/l')] += 1
                            stats['nlegal_max'] = max(stats['nlegal_max'], len(state['legal_actions']))
                    state, pid = env.step(a)
                    emit(1, a, env, state)
                for seat in range(P):                           # env.py:161-164 per-seat terminal states
                    emit(2, seat, env, env.get_state(seat))
                emit(3, 0, env, None, env.get_payoffs())
            if kind == 3 and reshuffles['n'] < 2:               # want >= 2 replace_deck calls in the episode
                for k in R:
                    del R[k][mark[k]:]
                del pay_rec[mark_pay:], pay[mark_pay:]
                continue
            stats['reshuffles'] += reshuffles['n']
            tapes.append(np.asarray(rec.tape, np.uint8))
            seeds.append(-1 if scripted else seed); kinds.append(kind); eps.append(episodes)
            rec_off.append(len(R['kind']))
            made += 1
    off = np.zeros(len(tapes) + 1, np.int64)
    off[1:] = np.cumsum([len(t) for t in tapes])
    out = dict(
        game=np.array(game), num_players=P, num_actions=A,
        slot_seed=np.asarray(seeds, np.int64), slot_kind=np.asarray(kinds, np.uint8), slot_episodes=np.asarray(eps, np.int32),
        tape=np.concatenate(tapes), tape_off=off, rec_off=np.asarray(rec_off, np.int64),
        rec_kind=np.asarray(R['kind'], np.uint8), rec_arg=np.asarray(R['arg'], np.int16),
        rec_player=np.asarray(R['player'], np.int8), rec_done=np.asarray(R['done'], np.uint8),
        rec_obs_dim=np.asarray(R['dim'], np.int16), rec_obs_hash=np.asarray(R['oh'], np.uint32),
        rec_legal_hash=np.asarray(R['lh'], np.uint32), rec_nlegal=np.asarray(R['nl'], np.uint16),
        pay_rec=np.asarray(pay_rec, np.int64), pay=np.stack(pay).astype(np.float32))
    assert np.array_equal(out['pay'].astype(np.float64), np.stack(pay))
    path = os.path.join(out_dir, 'deep_' + game.replace('-', '_') + '.npz')
    np.savez_compressed(path, **out)
    print('%-16s slots=%d episodes=%d records=%d steps=%d tape=%d B -> %s (%.0f KiB, %.0f s)' % (
        game, len(tapes), int(np.sum(eps)), len(R['kind']), int(np.sum(out['rec_kind'] == 1)), off[-1], path,
        os.path.getsize(path) / 1024, time.time() - t_start))
    if stats:
        print('   ', dict(stats))


if __name__ == '__main__':
    rl = import_reference()
    for g in (sys.argv[1:] or list(PLAN)):
        record_game(rl, g, HERE)
