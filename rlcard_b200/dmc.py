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


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def legal_ids(env, mask=None, max_ids=None):
    """Legal action ids of every env in ascending order: (ids int32 [N, max_ids] padded with -1, count int32 [N]).
    ``mask`` defaults to the env's current mask; ``max_ids`` to num_actions, for doudizhu to 1024 (the 20-card hand
    333444555666 + eight singles has 519 legal leads, the widest set found; tests/golden/deep_doudizhu.npz holds it).
    Raises if a list was cut: a truncated list would silently hide the highest ids (bombs, rocket) from the policy."""
    mask = env.mask if mask is None else mask
    if max_ids is None:
        max_ids = 1024 if env.name == 'doudizhu' else env.num_actions
    n = mask.shape[0]
    ids = torch.empty((n, max_ids), dtype=torch.int32, device=env.device)
    count = torch.empty(n, dtype=torch.int32, device=env.device)
    with torch.cuda.device(env.device):
        check(lib().rlc_legal_ids(env.gid, C.c_void_p(mask.data_ptr()), n, int(max_ids), C.c_void_p(ids.data_ptr()),
                                  C.c_void_p(count.data_ptr()), _stream(env.device)))
    widest = int(count.max().item()) if n else 0
    if widest > max_ids:
        raise ValueError('legal_ids: a state has %d legal actions, more than max_ids=%d' % (widest, max_ids))
    return ids, count


def action_features(env, ids):
    """env.get_action_feature for a tensor of action ids (any shape; -1 -> zeros): int8 [..., F]."""
    F = 54 if env.name == 'doudizhu' else env.num_actions
    flat = ids.to(device=env.device, dtype=torch.int32).contiguous().view(-1)
    out = torch.empty((flat.numel(), F), dtype=torch.int8, device=env.device)
    with torch.cuda.device(env.device):
        check(lib().rlc_action_features(env.gid, C.c_void_p(flat.data_ptr()), flat.numel(), C.c_void_p(out.data_ptr()), _stream(env.device)))
    return out.view(*ids.shape, F)


class DMCPolicy:
    """The DMC agents' action choice (dmc_agent/model.py:60-110) for all envs at once: every legal action of every env
    is scored by ``nets[position](obs, action_features)`` (the reference repeats the obs row per legal action), the
    best one is taken, or, with probability ``exp_epsilon``, a uniformly random legal one.  ``nets[p]`` maps
    (float32 [M, obs_dim_p], float32 [M, F]) -> values [M] or [M, 1]."""

    def __init__(self, env, nets, exp_epsilon=0.01, generator=None):
        self.env, self.nets, self.eps, self.gen = env, nets, float(exp_epsilon), generator

    def __call__(self, obs, mask, cur_player):
        env = self.env
        ids, count = legal_ids(env, mask)
        N, K = ids.shape
        valid = ids >= 0
        feats = action_features(env, ids)                                   # [N, K, F]
        values = torch.full((N, K), float('-inf'), device=env.device)
        for p in range(env.num_players):
            sel = (cur_player == p).nonzero(as_tuple=True)[0]
            if sel.numel() == 0:
                continue
            v = valid[sel]
            rows = v.nonzero(as_tuple=True)
            o = obs[sel][:, :env.obs_dims[p]].float()[rows[0]]
            val = self.nets[p](o, feats[sel][rows].float()).reshape(-1).float()
            block = torch.full((sel.numel(), K), float('-inf'), device=env.device)
            block[rows] = val
            values[sel] = block
        best = values.argmax(dim=1)                                         # first maximum = lowest id, like np.argmax
        rnd = (torch.rand(N, device=env.device, generator=self.gen) * count.clamp(min=1)).long().clamp(max=K - 1)
        explore = torch.rand(N, device=env.device, generator=self.gen) < self.eps
        pick = torch.where(explore, rnd, best)
        return ids.gather(1, pick.unsqueeze(1)).squeeze(1)


class DMCCollector:
    def __init__(self, env, pool_rows, open_capacity=None):
        """env: VecEnv.  pool_rows: capacity of each position's row pool.  open_capacity: bound on the number of
        decisions (all seats) of one episode -- the per-env store for episodes that span rollout windows."""
        self.env = env
        dev, N, P = env.device, env.num_envs, env.num_players
        self.F = 54 if env.name == 'doudizhu' else env.num_actions
        if open_capacity is None:
            open_capacity = {'blackjack': 16, 'leduc-holdem': 16, 'limit-holdem': 32, 'no-limit-holdem': 32, 'doudizhu': 192,
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
            # rows beyond a full pool were dropped (their slots were still counted): clamp so that the collector stays
            # usable after the caller has handled the error (drain with get_batch or clear())
            self.count.clamp_(max=self.pool_rows)
            self.overflow.zero_()
            raise RuntimeError('DMC pools overflowed (flag %d: 1 = a position pool was full, rows dropped; 2 = the open store '
                               'was too small, that episode was dropped whole)' % of)
        return [int(x) for x in self.count[:self.env.num_players].tolist()]

    def clear(self):
        """Forget every waiting row and every open episode (the envs keep running; decisions of episodes already under
        way are dropped until their episode ends)."""
        self.count.zero_(); self.overflow.zero_()
        self.open_len.fill_(-1)

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
