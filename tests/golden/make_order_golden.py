#!/usr/bin/env python
"""Golden fixture for the INSERTION ORDER of state['legal_actions'] (build container only; needs /root/reference).

``RandomAgent.step`` is ``np.random.choice(list(state['legal_actions'].keys()))`` (rlcard/agents/random_agent.py:27): the
draw indexes the OrderedDict, so a draw-exact replay of examples/run_random.py needs the reference's order, not just the
set.  UNO orders by first occurrence in the hand list (envs/uno.py:47-50 over games/uno/round.py:96-135), Scout lists the
multi-card plays, then the singles, then the scouts (games/scout/round.py:225-260).  This script plays
``env.run(is_training=True)`` with RandomAgents under a seeded global np.random and stores, per decision, the ordered id
list and the action taken, plus the payoffs of every episode.

Output: tests/golden/legal_order.npz.   Usage: python tests/golden/make_order_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

GAMES = {'uno': 30, 'scout': 6, 'leduc-holdem': 40, 'limit-holdem': 40, 'blackjack': 40, 'no-limit-holdem': 40, 'doudizhu': 0}
ENV_SEED, NP_SEED = 4242, 777


def main():
    rlcard = import_reference()
    from rlcard.agents import RandomAgent
    out = {}
    for game, episodes in GAMES.items():
        if not episodes:
            continue
        env = rlcard.make(game, config={'seed': ENV_SEED})
        ids, off, acts, pays = [], [0], [], []

        class Agent(RandomAgent):
            def step(self, state):
                keys = list(state['legal_actions'].keys())
                ids.extend(keys); off.append(len(ids))
                a = int(RandomAgent.step(state))
                acts.append(a)
                return a
        env.set_agents([Agent(num_actions=env.num_actions) for _ in range(env.num_players)])
        np.random.seed(NP_SEED)
        for _ in range(episodes):
            _, payoffs = env.run(is_training=True)
            pays.append(np.asarray(payoffs, np.float64))
        k = game.replace('-', '_')
        out[k + '_ids'] = np.asarray(ids, np.int16); out[k + '_off'] = np.asarray(off, np.int32)
        out[k + '_actions'] = np.asarray(acts, np.int16); out[k + '_payoffs'] = np.stack(pays)
        unsorted = sum(1 for i in range(len(acts)) if list(ids[off[i]:off[i + 1]]) != sorted(ids[off[i]:off[i + 1]]))
        print('%-16s episodes=%d decisions=%d of which %d list their ids in non-ascending order' % (game, episodes, len(acts), unsorted))
    out['env_seed'] = np.int64(ENV_SEED); out['np_seed'] = np.int64(NP_SEED)
    path = os.path.join(HERE, 'legal_order.npz')
    np.savez_compressed(path, **out)
    print('-> %s (%.1f KiB)' % (path, os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
