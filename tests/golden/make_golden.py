#!/usr/bin/env python
"""Generate golden replay fixtures from the LIVE reference engine.

Run in the build container only (needs /root/reference); the GPU box never
runs this.  It copies ``/root/reference/rlcard`` to a scratch directory (the
DouDizhu engine unzips its rule tables into its own package directory on
import, and /root/reference is read-only), adds a ``termcolor`` shim, wraps
``env.np_random`` with a recording proxy and plays seeded random episodes.

Output: ``tests/golden/<game>.npz`` with, per recorded *env slot* (one reference
``Env`` object; its episodes are played back to back because some engines leak
state across ``reset()``):

  slot_seed[n]          the ``seed`` given to ``rlcard.make``
  tape, tape_off        u8 chance tape: every bounded draw the reference RNG
                        made, as ``below(n)`` outcomes in consumption order.
                        shuffle(x) -> len(x)-1 draws (the reverse Fisher-Yates
                        swap indices j_i, i = n-1 .. 1, recovered from the
                        pre/post lists by object identity); randint(lo,hi) ->
                        one draw ``v-lo``; choice(n | list) -> one index draw.
  rec_*                 flat record stream (R records):
     rec_slot  slot index
     rec_kind  0 reset state, 1 state after step(rec_arg), 2 get_state(seat=rec_arg)
               view, 3 payoffs (rec_payoffs valid)
     rec_player current player after the event, rec_done is_over()
     rec_obs   [R, D] obs of the returned state (zero padded to the max seat dim)
     rec_obs_dim
     rec_legal [R, ceil(A/8)] packed bits (np.packbits, little bitorder)
     rec_payoffs [R, P]

Usage:  python tests/golden/make_golden.py [game ...]
"""
import os
import random
import shutil
import sys
import tempfile

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    scratch = os.path.join(tempfile.gettempdir(), 'rlcard_ref_copy')
    if not os.path.exists(os.path.join(scratch, 'rlcard')):
        os.makedirs(scratch, exist_ok=True)
        shutil.copytree(os.path.join(REF, 'rlcard'), os.path.join(scratch, 'rlcard'))
        os.makedirs(os.path.join(scratch, 'termcolor'), exist_ok=True)
        with open(os.path.join(scratch, 'termcolor', '__init__.py'), 'w') as f:
            f.write('def colored(s, *a, **k):\n    return s\n')
    sys.path.insert(0, scratch)
    import rlcard  # noqa
    return rlcard


class TapeRecorder:
    """Proxy for np.random.RandomState that logs bounded draws (see module doc)."""

    def __init__(self, rng):
        self.rng = rng
        self.tape = []

    def shuffle(self, x):
        pre = list(x)
        self.rng.shuffle(x)
        post = list(x)
        cur = pre
        for i in range(len(cur) - 1, 0, -1):
            tgt = post[i]
            j = next(k for k in range(i + 1) if cur[k] is tgt)
            cur[i], cur[j] = cur[j], cur[i]
            self.tape.append(j)
        assert all(a is b for a, b in zip(cur, post))

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        v = self.rng.randint(low, high)
        self.tape.append(int(v) - low)
        return v

    def choice(self, a):
        if isinstance(a, (int, np.integer)):
            v = self.rng.choice(a)
            self.tape.append(int(v))
            return v
        a = list(a)
        idx = self.rng.choice(len(a))   # same stream use as choice(list): randint(0, len)
        self.tape.append(int(idx))
        return a[idx]


GAMES = {
    # game: (slots, episodes per slot, prob of a raw random (maybe illegal) id, seat-view prob)
    'blackjack': (24, 40, 0.0, 1.0),
    'leduc-holdem': (24, 40, 0.15, 1.0),
    'limit-holdem': (24, 40, 0.15, 1.0),
    'uno': (12, 6, 0.0, 0.3),
    'doudizhu': (8, 3, 0.0, 0.25),
    'scout': (8, 2, 0.0, 0.2),
    'no-limit-holdem': (24, 40, 0.0, 1.0),
    # rare path: single long UNO episodes in which the draw pile runs out and UnoRound.replace_deck
    # (round.py:155-160) reshuffles the played pile back in; seeds are scanned until `slots` such episodes are found
    'uno-reshuffle': (6, 1, 0.0, 0.3),
}


def record_game(rlcard, game, out_dir):
    slots, episodes, p_raw, p_view = GAMES[game]
    fixture_name, want_reshuffle = game, game == 'uno-reshuffle'
    reshuffles = {'n': 0}
    if want_reshuffle:
        game = 'uno'
        from rlcard.games.uno.round import UnoRound
        if not hasattr(UnoRound, '_orig_replace_deck'):
            UnoRound._orig_replace_deck = UnoRound.replace_deck

            def counted(self):
                reshuffles['n'] += 1
                return UnoRound._orig_replace_deck(self)
            UnoRound.replace_deck = counted
    probe = rlcard.make(game, config={'seed': 0})
    P, A = probe.num_players, probe.num_actions
    D = max(int(np.prod(s)) for s in probe.state_shape)
    obs_dtype = np.float32 if game == 'scout' else np.int16
    R = dict(slot=[], kind=[], arg=[], player=[], done=[], obs=[], obs_dim=[], legal=[], payoffs=[])
    tapes, seeds = [], []

    def emit(slot, kind, arg, env, state, payoffs=None):
        R['slot'].append(slot); R['kind'].append(kind); R['arg'].append(arg)
        R['player'].append(env.get_player_id()); R['done'].append(int(env.is_over()))
        obs = np.zeros(D, obs_dtype); legal = np.zeros(A, np.uint8); od = 0
        if state is not None:
            o = np.asarray(state['obs']).reshape(-1)
            assert np.array_equal(o.astype(obs_dtype).astype(o.dtype), o)
            od = o.size
            obs[:od] = o
            ids = list(state['legal_actions'].keys())
            legal[ids] = 1
        R['obs'].append(obs); R['obs_dim'].append(od)
        R['legal'].append(np.packbits(legal, bitorder='little'))
        R['payoffs'].append(np.zeros(P) if payoffs is None else np.asarray(payoffs, np.float64))

    slot, candidate = 0, 0
    while slot < slots:
        seed = 1000 * (list(GAMES).index(fixture_name) + 1) + candidate
        candidate += 1
        mark = {k: len(v) for k, v in R.items()}
        reshuffles['n'] = 0
        env = rlcard.make(game, config={'seed': seed})
        rec = TapeRecorder(env.np_random)
        env.np_random = rec
        env.game.np_random = rec
        pol = random.Random(seed * 7919 + 1)
        seeds.append(seed)

        def views(always):
            if always or pol.random() < p_view:
                for seat in range(P):
                    try:
                        st = env.get_state(seat)
                    except UnboundLocalError:      # doudizhu peasants before the landlord has acted (envs/doudizhu.py:62-65)
                        continue
                    emit(slot, 2, seat, env, st)

        for ep in range(episodes):
            state, pid = env.reset()
            emit(slot, 0, 0, env, state)
            views(False)
            while not env.is_over():
                legal = list(state['legal_actions'].keys())
                if p_raw and pol.random() < p_raw:
                    a = pol.randrange(A)
                else:
                    a = legal[pol.randrange(len(legal))]
                state, pid = env.step(a)
                emit(slot, 1, a, env, state)
                views(env.is_over())
            emit(slot, 3, 0, env, None, env.get_payoffs())
        if want_reshuffle and reshuffles['n'] == 0:          # no reshuffle in this episode: drop it, try the next seed
            for k in R:
                del R[k][mark[k]:]
            seeds.pop()
            continue
        tapes.append(np.asarray(rec.tape, np.uint8))
        slot += 1

    off = np.zeros(slots + 1, np.int64)
    off[1:] = np.cumsum([len(t) for t in tapes])
    out = dict(
        game=np.array(game), num_players=P, num_actions=A, obs_max=D,
        slot_seed=np.asarray(seeds, np.int64), tape=np.concatenate(tapes), tape_off=off,
        rec_slot=np.asarray(R['slot'], np.int32), rec_kind=np.asarray(R['kind'], np.uint8),
        rec_arg=np.asarray(R['arg'], np.int32), rec_player=np.asarray(R['player'], np.int32),
        rec_done=np.asarray(R['done'], np.uint8), rec_obs=np.stack(R['obs']),
        rec_obs_dim=np.asarray(R['obs_dim'], np.int32), rec_legal=np.stack(R['legal']),
        rec_payoffs=np.stack(R['payoffs']))
    path = os.path.join(out_dir, fixture_name.replace('-', '_') + '.npz')
    np.savez_compressed(path, **out)
    print('%-14s slots=%d records=%d tape=%d bytes -> %s (%.1f KiB)' % (
        fixture_name, slots, len(R['slot']), off[-1], path, os.path.getsize(path) / 1024))


if __name__ == '__main__':
    rl = import_reference()
    for g in (sys.argv[1:] or list(GAMES)):
        record_game(rl, g, HERE)
