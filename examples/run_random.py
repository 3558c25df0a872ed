"""examples/run_random.py of the reference on the B200 simulator: one seeded episode through the single-env
facade (same trajectory as rlcard.make(env, {'seed': 42}) would deal), then the same loop for 65 536 envs at once."""
import argparse
import os
import sys
import pprint
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere in the checkout
import rlcard_b200
from rlcard_b200 import RandomAgent


def run(args):
    env = rlcard_b200.make(args.env, config={'seed': 42})
    np.random.seed(42)
    agent = RandomAgent(num_actions=env.num_actions)
    env.set_agents([agent for _ in range(env.num_players)])
    trajectories, payoffs = env.run(is_training=False)
    print('payoffs:', payoffs)
    print('steps of seat 0:', len(trajectories[0]) // 2)
    print('sample legal_actions:')
    pprint.pprint(trajectories[0][0]['raw_legal_actions'])

    vec = rlcard_b200.VecEnv(args.env, args.num_envs, seed=42)
    vec.reset()
    traj = vec.rollout_random(args.steps)                 # warm-up + first window
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    traj = vec.rollout_random(args.steps, out=traj)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    done = traj['done'].bool()
    print('%d envs x %d steps in %.2f ms: %.3g env-steps/s, %d episodes, mean payoff per seat %s' % (
        args.num_envs, args.steps, dt * 1e3, args.num_envs * args.steps / dt, int(done.sum()),
        (traj['payoffs'][done].mean(0)).tolist()))


if __name__ == '__main__':
    ap = argparse.ArgumentParser('Random example on the B200 simulator')
    ap.add_argument('--env', default='leduc-holdem', choices=sorted(rlcard_b200.GAME_IDS))
    ap.add_argument('--num-envs', type=int, default=65536)
    ap.add_argument('--steps', type=int, default=128)
    run(ap.parse_args())
