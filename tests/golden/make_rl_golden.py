#!/usr/bin/env python
"""Golden fixture for run_rl.py's data path, from the LIVE reference (build container only): episodes of
``env.run(is_training=True)`` with recorded random agents, reorganised by the reference's own ``reorganize``
(rlcard/utils/utils.py:153-179).  Stored per game: chance tape, the action ids in order, and per seat the
transitions (state obs, action, reward, next_state obs, legal ids of next_state, done).

Output: tests/golden/rl_<game>.npz.   Usage: python tests/golden/make_rl_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import TapeRecorder, import_reference  # noqa: E402

GAMES = {'leduc-holdem': 12, 'limit-holdem': 8, 'uno': 3, 'doudizhu': 2, 'blackjack': 10, 'scout': 1, 'no-limit-holdem': 10}


def main():
    rlcard = import_reference()
    from rlcard.utils.utils import reorganize
    for game, episodes in GAMES.items():
        if len(sys.argv) > 1 and game not in sys.argv[1:]:
            continue
        env = rlcard.make(game, config={'seed': 21})
        rec = TapeRecorder(env.np_random)
        env.np_random = rec
        env.game.np_random = rec
        P, A = env.num_players, env.num_actions
        rng = np.random.RandomState(99)
        actions = []

        class Agent:
            use_raw = False

            def step(self, state):
                ids = list(state['legal_actions'].keys())
                a = int(ids[rng.randint(len(ids))])
                actions.append(a)
                return a
        env.set_agents([Agent() for _ in range(P)])
        D = max(int(np.prod(s)) for s in env.state_shape)
        rows = [[] for _ in range(P)]
        for _ in range(episodes):
            traj, payoffs = env.run(is_training=True)
            for p, ts in enumerate(reorganize(traj, payoffs)):
                rows[p] += ts
        out = dict(game=np.array(game), tape=np.asarray(rec.tape, np.uint8), actions=np.asarray(actions, np.int32),
                   num_players=np.int32(P), episodes=np.int32(episodes))
        dt = np.float32 if game == 'scout' else np.int16
        for p in range(P):
            n = len(rows[p])
            st = np.zeros((n, D), dt); nx = np.zeros((n, D), dt); lg = np.zeros((n, A), np.uint8)
            for i, (s, a, r, s2, d) in enumerate(rows[p]):
                o, o2 = np.asarray(s['obs']).reshape(-1), np.asarray(s2['obs']).reshape(-1)
                st[i, :o.size] = o; nx[i, :o2.size] = o2
                lg[i, list(s2['legal_actions'].keys())] = 1
            out['state_%d' % p] = st; out['next_state_%d' % p] = nx
            out['next_legal_%d' % p] = np.packbits(lg, axis=1, bitorder='little')
            out['action_%d' % p] = np.asarray([t[1] for t in rows[p]], np.int32)
            out['reward_%d' % p] = np.asarray([t[2] for t in rows[p]], np.float64)
            out['done_%d' % p] = np.asarray([t[4] for t in rows[p]], np.uint8)
        path = os.path.join(HERE, 'rl_%s.npz' % game.replace('-', '_'))
        np.savez_compressed(path, **out)
        print('%-13s episodes=%d transitions=%s -> %s (%.1f KiB)' % (game, episodes, [len(r) for r in rows], path, os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
