"""DMC actor data path on device: the reference's ``act`` loop (rlcard/agents/dmc_agent/utils.py:97-155) and
``get_batch`` (:33-50) over a VecEnv.

The reference runs one Python env per actor process and appends, per position, (done, episode_return, target,
state int8, action int8) rows that the learner consumes in ``[T, B]`` batches.  ``DMCCollector`` keeps those row
pools in HBM: a rollout window of the VecEnv is folded into them by ``rlc_dmc_collect`` (one warp per env, see
csrc/tu_dmc.cu), and ``get_batch(position, T, B)`` hands out ``[T, B, ...]`` tensors with the reference's keys
and dtypes.
"""
import ctypes as C

import torch

from ._lib import DTYPE_F32, DTYPE_U8, RlcDmcBuffers, RlcTrajectory, check, lib


class DMCCollector:
    def __init__(self, env, pool_rows, open_capacity=None):
        """env: VecEnv.  pool_rows: capacity of each position's row pool.  open_capacity: bound on the number of
        decisions (all seats) of one episode -- the per-env store for episodes that span rollout windows."""
        self.env = env
        dev, N, P = env.device, env.num_envs, env.num_players
        self.F = 54 if env.name == 'doudizhu' else env.num_actions
        if open_capacity is None:
            open_capacity = {'blackjack': 16, 'leduc-holdem': 16, 'limit-holdem': 32, 'doudizhu': 192,
                             'uno': 1024, 'scout': 1024}[env.name]
        self.open_capacity, self.pool_rows = int(open_capacity), int(pool_rows)
        odt = env.obs_dtype
        self.open_obs = torch.zeros((N, self.open_capacity, env.obs_stride), dtype=odt, device=dev)
        self.open_action = torch.zeros((N, self.open_capacity), dtype=torch.int32, device=dev)
        self.open_player = torch.zeros((N, self.open_capacity), dtype=torch.int8, device=dev)
        self.open_len = torch.zeros(N, dtype=torch.int32, device=dev)
        self.state = [torch.zeros((self.pool_rows, env.obs_stride), dtype=odt, device=dev) for _ in range(P)]
        self.action = [torch.zeros((self.pool_rows, self.F), dtype=torch.int8, device=dev) for _ in range(P)]
        self.target = [torch.zeros(self.pool_rows, dtype=torch.float32, device=dev) for _ in range(P)]
        self.episode_return = [torch.zeros(self.pool_rows, dtype=torch.float32, device=dev) for _ in range(P)]
        self.done = [torch.zeros(self.pool_rows, dtype=torch.uint8, device=dev) for _ in range(P)]
        self.count = torch.zeros(4, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        b = RlcDmcBuffers()
        b.open_obs, b.open_action = self.open_obs.data_ptr(), self.open_action.data_ptr()
        b.open_player, b.open_len, b.open_capacity = self.open_player.data_ptr(), self.open_len.data_ptr(), self.open_capacity
        for p in range(P):
            b.out_state[p], b.out_action[p] = self.state[p].data_ptr(), self.action[p].data_ptr()
            b.out_target[p], b.out_episode_return[p] = self.target[p].data_ptr(), self.episode_return[p].data_ptr()
            b.out_done[p] = self.done[p].data_ptr()
        b.out_count, b.out_capacity, b.overflow = self.count.data_ptr(), self.pool_rows, self.overflow.data_ptr()
        self._b = b

    def add(self, traj, T=None):
        """Fold a trajectory window (dict of [T, N, ...] device tensors as produced by VecEnv.rollout_random or
        assembled from reset()/step()) into the pools."""
        env = self.env
        T = int(traj['action'].shape[0] if T is None else T)
        tr = RlcTrajectory()
        for k in ('obs', 'action', 'player', 'done', 'payoffs'):
            assert traj[k].is_contiguous()
            setattr(tr, k, traj[k].data_ptr())
        with torch.cuda.device(env.device):
            check(lib().rlc_dmc_collect(env.gid, C.byref(tr), DTYPE_F32 if env.obs_dtype == torch.float32 else DTYPE_U8,
                                        T, env.num_envs, C.byref(self._b),
                                        C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)))

    def collect_random(self, T, traj=None):
        """One rollout window with the on-device random agents (the actors' exploration at exp_epsilon = 1)."""
        traj = self.env.rollout_random(T, out=traj)
        self.add(traj, T)
        return traj

    def sizes(self):
        of = int(self.overflow.item())
        if of:
            raise RuntimeError('DMC pools overflowed (flag %d: 1 = a position pool was full, 2 = open store too small)' % of)
        return [int(x) for x in self.count[:self.env.num_players].tolist()]

    def get_batch(self, position, T, B):
        """The learner's batch for one position (dmc_agent/utils.py:33-50): dict of [T, B, ...] tensors with keys
        done (bool), episode_return, target (float32), state (int8 [.., obs_dim of the position]), action (int8).
        Returns None while fewer than T * B rows are waiting.  Column b holds T consecutive rows of the pool."""
        p, need = int(position), int(T) * int(B)
        have = self.sizes()[p]
        if have < need:
            return None
        d = self.env.obs_dims[p]
        tb = lambda x: x[:need].reshape(B, T, *x.shape[1:]).transpose(0, 1).contiguous()
        batch = {'done': tb(self.done[p]).bool(), 'episode_return': tb(self.episode_return[p]), 'target': tb(self.target[p]),
                 'state': tb(self.state[p][:, :d]).to(torch.int8), 'action': tb(self.action[p])}
        rest = have - need                                      # slide the unread rows to the front of the pool
        for buf in (self.state[p], self.action[p], self.target[p], self.episode_return[p], self.done[p]):
            buf[:rest] = buf[need:have].clone()
        self.count[p] = rest
        return batch
