"""The data side of examples/run_dmc.py on the B200 simulator: actors with (random-init) DMC-style nets drive a
VecEnv, rlc_dmc_collect fills the per-position pools, get_batch hands [T, B] batches to a learner step."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere in the checkout
import rlcard_b200
from rlcard_b200.dmc import DMCCollector, DMCPolicy


class Net(torch.nn.Module):                      # the shape of dmc_agent/model.py DMCNet: MLP over concat(obs, action feature)
    def __init__(self, obs_dim, feat_dim, width=512):
        super().__init__()
        self.f = torch.nn.Sequential(torch.nn.Linear(obs_dim + feat_dim, width), torch.nn.ReLU(),
                                     torch.nn.Linear(width, width), torch.nn.ReLU(), torch.nn.Linear(width, 1))

    def forward(self, obs, actions):
        return self.f(torch.cat([obs, actions], dim=-1)).flatten()


def main(args):
    env = rlcard_b200.VecEnv(args.env, args.num_envs, seed=0)
    env.reset()
    F = 54 if args.env == 'doudizhu' else env.num_actions
    nets = [Net(d, F).cuda() for d in env.obs_dims]
    opts = [torch.optim.RMSprop(n.parameters(), lr=1e-4) for n in nets]
    col = DMCCollector(env, pool_rows=args.num_envs * args.window * 2)
    policy = DMCPolicy(env, [lambda o, a, n=n: n(o, a).detach() for n in nets], exp_epsilon=0.01)
    keys = ('obs', 'action', 'player', 'done', 'payoffs')
    for it in range(args.iterations):
        traj = {k: [] for k in keys}
        with torch.no_grad():
            for _ in range(args.window):
                a = policy(env.obs, env.mask, env.cur_player)
                traj['obs'].append(env.obs.clone()); traj['player'].append(env.cur_player.clone()); traj['action'].append(a.int())
                env.step(a)
                traj['done'].append(env.done.clone()); traj['payoffs'].append(env.payoffs.clone())
        col.add({k: torch.stack(v).contiguous() for k, v in traj.items()})
        for p, (net, opt) in enumerate(zip(nets, opts)):
            batch = col.get_batch(p, args.unroll, args.batch)
            if batch is None:
                continue
            col.count[p] = 0                             # this toy learner takes one batch per iteration and drops the rest
            state = batch['state'].flatten(0, 1).float(); action = batch['action'].flatten(0, 1).float()
            loss = ((net(state, action) - batch['target'].flatten()) ** 2).mean()        # trainer.py:42-44 compute_loss
            opt.zero_grad(); loss.backward(); opt.step()
            ret = batch['episode_return'][batch['done']]
            print('iter %d position %d loss %.4f mean_episode_return %.3f waiting rows %d' % (
                it, p, float(loss), float(ret.mean()) if ret.numel() else float('nan'), col.sizes()[p]))
    env.check_errors()


if __name__ == '__main__':
    ap = argparse.ArgumentParser('DMC data path example on the B200 simulator')
    ap.add_argument('--env', default='doudizhu', choices=sorted(rlcard_b200.GAME_IDS))
    ap.add_argument('--num-envs', type=int, default=1024)
    ap.add_argument('--window', type=int, default=32)
    ap.add_argument('--unroll', type=int, default=100)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--iterations', type=int, default=4)
    main(ap.parse_args())
