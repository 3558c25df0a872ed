#!/usr/bin/env python
"""Golden fixture for the DMC actor data path, from the LIVE reference (build container only).

Runs the reference's own ``act`` loop (rlcard/agents/dmc_agent/utils.py:97-155, loaded by file path so that
the package's trainer / git imports are not needed) on a reference env with uniform-random agents, with the
same recording proxy as make_golden.py on the env's RandomState and a recorder on the agents' choices.  The
loop is stopped after K buffers per position through the queue stand-ins.  Stored per game:
chance tape, the action ids in the order they were taken, and every buffer the actor filled
(done / episode_return / target / state / action, [K, T, ...] per position).

Output: tests/golden/dmc_<game>.npz.   Usage: python tests/golden/make_dmc_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, TapeRecorder, import_reference  # noqa: E402

GAMES = {'doudizhu': (12, 3), 'leduc-holdem': (10, 4), 'uno': (16, 3), 'limit-holdem': (8, 3), 'scout': (16, 3)}   # game: (T, K buffers per position)


class StopActor(KeyboardInterrupt):
    pass


def record(rlcard, act, game, T, K):
    env = rlcard.make(game, config={'seed': 0})
    P = env.num_players
    rec_holder, actions = {}, []
    orig_seed = env.seed

    def seed(s):                                   # act() reseeds the env first (utils.py:111): record from there on
        orig_seed(s)
        rec = TapeRecorder(env.np_random)
        env.np_random = rec
        env.game.np_random = rec
        rec_holder['rec'] = rec
    env.seed = seed

    class Agent:                                    # uniform over the legal ids like agents/random_agent.py:17-27
        use_raw = False

        def __init__(self, rng):
            self.rng = rng

        def step(self, state):
            ids = list(state['legal_actions'].keys())
            a = int(ids[self.rng.randint(len(ids))])
            actions.append(a)
            return a

    class Model:
        def get_agents(self):
            rng = np.random.RandomState(12345)
            return [Agent(rng) for _ in range(P)]

    feat = 54 if game == 'doudizhu' else env.num_actions
    buffers = []
    for p in range(P):
        shape = tuple(env.state_shape[p])
        buffers.append(dict(
            done=[torch.zeros((T,), dtype=torch.bool) for _ in range(K)],
            episode_return=[torch.zeros((T,), dtype=torch.float32) for _ in range(K)],
            target=[torch.zeros((T,), dtype=torch.float32) for _ in range(K)],
            state=[torch.zeros((T,) + shape, dtype=torch.int8) for _ in range(K)],
            action=[torch.zeros((T, feat), dtype=torch.int8) for _ in range(K)]))

    class Free:
        exhausted = set()

        def __init__(self, p):
            self.next, self.p = 0, p

        def get(self):
            if self.next >= K:                      # this position has its K buffers: the actor skips it (utils.py:139-141)
                Free.exhausted.add(self.p)
                if len(Free.exhausted) == P:
                    raise StopActor()
                return None
            self.next += 1
            return self.next - 1

    class Full:
        def __init__(self):
            self.got = []

        def put(self, i):
            self.got.append(i)

    Free.exhausted = set()
    free, full = [Free(p) for p in range(P)], [Full() for _ in range(P)]
    act(7, 'cpu', T, free, full, Model(), buffers, env)              # returns on the KeyboardInterrupt subclass
    out = dict(game=np.array(game), T=np.int32(T), seed=np.int64(7), num_players=np.int32(P),
               tape=np.asarray(rec_holder['rec'].tape, np.uint8), actions=np.asarray(actions, np.int32))
    for p in range(P):
        k = len(full[p].got)
        out['filled_%d' % p] = np.int32(k)
        for key in ('done', 'episode_return', 'target', 'state', 'action'):
            arr = torch.stack([buffers[p][key][i] for i in full[p].got]).numpy() if k else np.zeros((0, T))
            out['%s_%d' % (key, p)] = arr.reshape(k, T, -1) if key in ('state', 'action') else arr
    path = os.path.join(HERE, 'dmc_%s.npz' % game.replace('-', '_'))
    np.savez_compressed(path, **out)
    print('%-13s T=%d filled=%s actions=%d tape=%d -> %s (%.1f KiB)' % (
        game, T, [int(out['filled_%d' % p]) for p in range(P)], len(actions), len(out['tape']), path,
        os.path.getsize(path) / 1024))


def main():
    rlcard = import_reference()
    spec = importlib.util.spec_from_file_location('ref_dmc_utils', os.path.join(REF, 'rlcard', 'agents', 'dmc_agent', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.log.setLevel(100)
    for game, (T, K) in GAMES.items():
        if len(sys.argv) > 1 and game not in sys.argv[1:]:
            continue
        record(rlcard, mod.act, game, T, K)


if __name__ == '__main__':
    main()
