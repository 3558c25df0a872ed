"""The data side of examples/run_rl.py on the B200 simulator: a Q-network plays seat 0 epsilon-greedily, the other seats
play randomly, TransitionCollector yields (state, action, reward, next_state, next legal mask, done) for a DQN update."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere in the checkout
import rlcard_b200
from rlcard_b200.rl import TransitionCollector


def main(args):
    env = rlcard_b200.VecEnv(args.env, args.num_envs, seed=0, auto_reset=False, terminal_obs=True)
    if env.mask_bitpacked:
        raise SystemExit('this toy DQN head scores a dense action vector: pick a game with a dense mask')
    col = TransitionCollector(env, pool_rows=args.num_envs * args.steps)
    col.reset()
    q = torch.nn.Sequential(torch.nn.Linear(env.obs_dims[0], 64), torch.nn.ReLU(), torch.nn.Linear(64, env.num_actions)).cuda()
    opt = torch.optim.Adam(q.parameters(), lr=5e-4)
    rnd = rlcard_b200.random_policy()

    def policy(obs, mask, cur):                          # dqn_agent.py:142-160 step, batched
        with torch.no_grad():
            qv = q(obs[:, :env.obs_dims[0]].float()).masked_fill(mask == 0, float('-inf'))
        greedy = qv.argmax(1).int()
        explore = torch.rand(obs.shape[0], device=obs.device) < args.epsilon
        return torch.where((cur == 0) & ~explore, greedy, rnd(obs, mask, cur))

    for it in range(args.iterations):
        col.run(policy, args.steps)
        b = col.pop(0)
        for seat in range(1, env.num_players):
            col.pop(seat)                                # only seat 0 learns (run_rl.py:84-86)
        if b['action'].numel() == 0:
            continue
        with torch.no_grad():                            # dqn_agent.py:197-233 train: masked double-free target
            nq = q(b['next_state'].float()).masked_fill(b['next_mask'] == 0, float('-inf')).max(1).values
            nq = torch.where(b['done'] | torch.isinf(nq), torch.zeros_like(nq), nq)
            target = b['reward'] + 0.99 * nq
        pred = q(b['state'].float()).gather(1, b['action'].long().unsqueeze(1)).squeeze(1)
        loss = torch.nn.functional.mse_loss(pred, target)
        opt.zero_grad(); loss.backward(); opt.step()
        print('iter %d: %d transitions, loss %.4f, mean reward of finished episodes %.3f' % (
            it, b['action'].numel(), float(loss), float(b['reward'][b['done']].mean())))
    env.check_errors()


if __name__ == '__main__':
    ap = argparse.ArgumentParser('DQN data path example on the B200 simulator')
    ap.add_argument('--env', default='leduc-holdem', choices=sorted(rlcard_b200.GAME_IDS))
    ap.add_argument('--num-envs', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=32)
    ap.add_argument('--epsilon', type=float, default=0.1)
    ap.add_argument('--iterations', type=int, default=5)
    main(ap.parse_args())
