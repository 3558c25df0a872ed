"""run_rl.py's data path on device: ``env.run(is_training=True)`` + ``reorganize`` (rlcard/utils/utils.py:153-179)
as a per-step stream over a VecEnv.

``TransitionCollector.step(actions)`` applies one action per env and appends, per seat, the transitions
``[state, action, reward, next_state, done]`` the reference would hand to ``agent.feed`` (agents/dqn_agent.py:127-140,
which also wants the legal actions of ``next_state``): each decision of a seat is paired with the seat's next
decision, the last one with its terminal view and payoff.  Rows live in per-seat pools in HBM.
"""
import ctypes as C

import torch

from ._lib import AUTO_RESET, DTYPE_F32, DTYPE_U8, TERMINAL_OBS, RlcRlBuffers, RlcTrajectory, check, lib


class TransitionCollector:
    def __init__(self, env, pool_rows, fused=False):
        """env: VecEnv created with ``terminal_obs=True`` (the per-seat terminal views are part of the transitions);
        ``fused=True`` -> only the fused-rollout path (``collect_fused``) is used and any VecEnv will do."""
        if env.terminal_obs is None and not fused:
            raise ValueError('TransitionCollector needs VecEnv(..., terminal_obs=True)')
        self.env = env
        self._win = None
        dev, N, P = env.device, env.num_envs, env.num_players
        self.pool_rows = int(pool_rows)
        odt = env.obs_dtype
        mshape = (env.mask_words,) if env.mask_bitpacked else (env.num_actions,)
        mdt = torch.int32 if env.mask_bitpacked else torch.uint8
        self.pend_obs = torch.zeros((N, P, env.obs_stride), dtype=odt, device=dev)
        self.pend_action = torch.zeros((N, P), dtype=torch.int32, device=dev)
        self.pend_valid = torch.zeros((N, P), dtype=torch.uint8, device=dev)
        mk = lambda shape, dt: [torch.zeros((self.pool_rows,) + shape, dtype=dt, device=dev) for _ in range(P)]
        self.state, self.next_state = mk((env.obs_stride,), odt), mk((env.obs_stride,), odt)
        self.next_mask = mk(mshape, mdt)
        self.action, self.reward, self.done = mk((), torch.int32), mk((), torch.float32), mk((), torch.uint8)
        self.count = torch.zeros(4, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        b = RlcRlBuffers()
        b.pend_obs, b.pend_action, b.pend_valid = self.pend_obs.data_ptr(), self.pend_action.data_ptr(), self.pend_valid.data_ptr()
        for p in range(P):
            b.out_state[p], b.out_action[p], b.out_reward[p] = self.state[p].data_ptr(), self.action[p].data_ptr(), self.reward[p].data_ptr()
            b.out_next_state[p], b.out_next_mask[p], b.out_done[p] = self.next_state[p].data_ptr(), self.next_mask[p].data_ptr(), self.done[p].data_ptr()
        b.out_count, b.out_capacity, b.overflow = self.count.data_ptr(), self.pool_rows, self.overflow.data_ptr()
        self._b = b

    def _feed(self, phase, actions):
        env = self.env
        with torch.cuda.device(env.device):
            check(lib().rlc_rl_feed(env.gid, phase, C.byref(env._buffers()), None if actions is None else C.c_void_p(actions.data_ptr()),
                                    env.num_envs, C.byref(self._b), C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)))

    def reset(self):
        self.pend_valid.zero_()
        return self.env.reset()

    def step(self, actions):
        """One env-step for every env (Env.step), transitions appended, finished envs dealt again.
        Returns (obs, mask, cur_player, done, payoffs) like VecEnv.step; obs/mask/cur_player describe the state the
        next action applies to (the fresh deal for the envs flagged done)."""
        env = self.env
        actions = actions.to(device=env.device, dtype=torch.int32).contiguous()
        self._feed(0, actions)
        env.step(actions, auto_reset=False)
        self._feed(1, None)
        done = env.done.clone()
        payoffs = env.payoffs.clone()
        env.reset(done)
        return env.obs, env.mask, env.cur_player, done, payoffs

    def run(self, policy, num_steps):
        """The run_rl.py loop body: ``policy(obs, mask, cur_player) -> actions`` drives every seat."""
        obs, mask, cur = self.env.obs, self.env.mask, self.env.cur_player
        for _ in range(num_steps):
            obs, mask, cur, _, _ = self.step(policy(obs, mask, cur))

    def collect_fused(self, T, actions=None, terminal_rows=None):
        """run_rl.py's data path on the FUSED rollout: one rlc_rollout_random launch of T env-steps per env (random
        agents, or the recorded ids in ``actions`` int32 [T, N]; auto reset) that also writes the per-seat terminal states
        into a pool, then one rlc_reorganize launch that folds the window into the per-seat transition pools.  Consecutive
        calls chain: a decision whose next state falls into the next window waits in pend_*.  Returns the window."""
        env = self.env
        rows = int(terminal_rows or env.num_envs * T)
        if self._win is None or self._win['action'].shape[0] != T or self._win['terminal_obs'].shape[0] != rows:
            self._win = env.alloc_terminal_pool(rows, T, env.alloc_trajectory(T))
        w = self._win
        w['terminal_count'].zero_()
        env.rollout_random(T, out=w, actions=actions)
        tr = RlcTrajectory()
        for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs', 'terminal_obs', 'terminal_mask', 'terminal_row'):
            setattr(tr, k, w[k].data_ptr())
        with torch.cuda.device(env.device):
            check(lib().rlc_reorganize(env.gid, C.byref(tr), DTYPE_F32 if env.obs_dtype == torch.float32 else DTYPE_U8, int(T),
                                       env.num_envs, C.byref(self._b), C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)))
        return w

    def sizes(self):
        if int(self.overflow.item()):
            raise RuntimeError('transition pools overflowed')
        return [int(x) for x in self.count[:self.env.num_players].tolist()]

    def pop(self, seat):
        """All transitions waiting for one seat as a dict of tensors (state, action, reward, next_state, next_mask, done),
        state rows cut to the seat's obs width; the pool is emptied."""
        p = int(seat)
        n, d = self.sizes()[p], self.env.obs_dims[p]
        out = {'state': self.state[p][:n, :d].clone(), 'action': self.action[p][:n].clone(), 'reward': self.reward[p][:n].clone(),
               'next_state': self.next_state[p][:n, :d].clone(), 'next_mask': self.next_mask[p][:n].clone(),
               'done': self.done[p][:n].bool()}
        self.count[p] = 0
        return out
