"""VecEnv: N independent card-game envs stepped by CUDA kernels through the C ABI.

Drop-in for the reference's rollout loops (``Env.run`` in rlcard/envs/env.py:120-169, the body of
examples/run_random.py:23-27, ``tournament`` in rlcard/utils/utils.py:200-225 and the DMC actor loop
rlcard/agents/dmc_agent/utils.py:121-137): instead of one Python ``Env`` per process, one ``VecEnv``
per GPU holds every env's packed state in HBM.  torch owns all memory; the library gets raw pointers.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (AUTO_RESET, CHANCE_MT19937, CHANCE_PHILOX, CHANCE_TAPE, DTYPE_F32, DTYPE_U8, GAME_IDS, STATE_ROWS,
                   TERMINAL_OBS, RlcBuffers, RlcTrajectory, check, game_info, lib)

_MODES = {'throughput': CHANCE_PHILOX, 'philox': CHANCE_PHILOX, 'replay': CHANCE_TAPE, 'tape': CHANCE_TAPE,
          'mt19937': CHANCE_MT19937}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_TABLES_ON = set()


def _upload_doudizhu_tables(L, device):
    """The constant action table (games/doudizhu/jsondata.zip in the reference, utils.py:14-38) goes to the
    device once; the library keeps it until process exit."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _TABLES_ON:
        return
    from . import doudizhu_table
    blob = doudizhu_table.build_blob()
    buf = (C.c_char * len(blob)).from_buffer_copy(blob)
    check(L.rlc_upload_tables(GAME_IDS['doudizhu'], idx, C.cast(buf, C.c_void_p), len(blob)))
    _TABLES_ON.add(idx)


def seed_words(seed):
    """rlcard/utils/seeding.py:33-113: sha512(str(seed))[:8] -> uint32 words given to RandomState.seed."""
    import hashlib
    import struct
    seed = int(seed) % 2 ** 64
    lo, hi = struct.unpack('2I', hashlib.sha512(str(seed).encode('utf8')).digest()[:8])
    return [lo, hi] if hi else [lo]


class VecEnv:
    """``num_envs`` environments of one game on one GPU.

    mode 'throughput': Philox4x32-10 chance keyed by (seed, env_id_base + i, env-step index).
    mode 'replay':     chance draws come from a uint8 tape per env (``set_tape``), i.e. the
                       reference's recorded np.random outcomes.
    mode 'mt19937':    np.random.RandomState on device, seeded per env like ``rlcard.make(seed=...)``.
    """

    def __init__(self, env_id, num_envs, device='cuda:0', seed=0, mode='throughput', obs_dtype=None,
                 env_id_base=0, auto_reset=True, terminal_obs=False, legal_order=False):
        if env_id not in GAME_IDS:
            raise ValueError('unknown env id %r (have %s)' % (env_id, sorted(GAME_IDS)))
        self.L = lib()                     # raises if the CUDA extension is missing
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.RlcError('rlcard_b200 runs on CUDA devices only (no CPU fallback)')
        self.name = env_id
        self.gid = GAME_IDS[env_id]
        info = game_info(self.gid)
        if info.state_words == 0:
            raise _lib.RlcError('game %r has no kernels in this build' % env_id)
        self.num_envs = int(num_envs)
        self.num_players = info.num_players
        self.num_actions = info.num_actions
        self.obs_dims = list(info.obs_dim)[:info.num_players]
        self.obs_stride = info.obs_stride
        self.mask_bitpacked = bool(info.mask_bitpacked)
        self.mask_words = info.mask_words
        self.state_words = info.state_words
        self.state_rows = info.state_layout == STATE_ROWS        # [N][state_words] (warp-per-env games) vs [state_words][N]
        self.chance = _MODES[mode]
        self.seed_value = int(seed)
        self.env_id_base = int(env_id_base)
        self.auto_reset = bool(auto_reset)
        self.terminal_obs_enabled = bool(terminal_obs)
        if obs_dtype is None:
            obs_dtype = torch.float32 if info.obs_native_dtype == DTYPE_F32 else torch.uint8
        if obs_dtype not in (torch.uint8, torch.float32):
            raise ValueError('obs_dtype must be torch.uint8 or torch.float32')
        self.obs_dtype = obs_dtype
        N, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            self.state = torch.zeros((N, self.state_words) if self.state_rows else (self.state_words, N),
                                     dtype=torch.int32, device=dev)
            # the five per-step outputs are views of ONE device buffer (256-byte aligned sections), so a host
            # consumer fetches a whole step with a single device-to-host copy (step_host)
            esz = 4 if obs_dtype == torch.float32 else 1
            msz = (self.mask_words * 4) if self.mask_bitpacked else self.num_actions
            sizes = [N * self.obs_stride * esz, N * msz, N * 4, N, N * self.num_players * 4,
                     N * min(self.num_actions, 1024) * 4 if legal_order else 0]
            offs, cur = [], 0
            for sz in sizes:
                offs.append(cur)
                cur += (sz + 255) & ~255
            self._out = torch.zeros(cur, dtype=torch.uint8, device=dev)
            sec = lambda k: self._out[offs[k]:offs[k] + sizes[k]]
            self._sections = (offs, sizes)
            self.obs = sec(0).view(obs_dtype).view(N, self.obs_stride)
            if self.mask_bitpacked:
                self.mask = sec(1).view(torch.int32).view(N, self.mask_words)
            else:
                self.mask = sec(1).view(N, self.num_actions)
            self.cur_player = sec(2).view(torch.int32)
            self.done = sec(3)
            self.payoffs = sec(4).view(torch.float32).view(N, self.num_players)
            self.err = torch.zeros(N, dtype=torch.int32, device=dev)
            self.terminal_obs = (torch.zeros((N, self.num_players, self.obs_stride), dtype=obs_dtype, device=dev)
                                 if terminal_obs else None)
            # legal ids in the reference's insertion order (rlc_buffers.legal_order), -1 padded; doudizhu rows are cut
            # at 1024 ids (widest legal set: 519)
            self.legal_order = sec(5).view(torch.int32).view(N, min(self.num_actions, 1024)) if legal_order else None
            if legal_order:
                self.legal_order.fill_(-1)
        self.tape = None
        self.tape_pos = None
        self.mt = None
        self._buf = None
        self.launches = 0
        if env_id == 'doudizhu':
            _upload_doudizhu_tables(self.L, self.device)
        elif env_id in ('leduc-holdem', 'limit-holdem'):   # betting-state table of the tabulated rollout: explicit one-time init
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            check(self.L.rlc_upload_tables(self.gid, idx, None, 0))

    # ------------------------------------------------------------------ chance sources
    def set_tape(self, tape):
        """tape: uint8 [num_envs, L] draw outcomes in consumption order (replay mode)."""
        tape = torch.as_tensor(np.ascontiguousarray(tape, dtype=np.uint8)).to(self.device)
        assert tape.dim() == 2 and tape.shape[0] == self.num_envs
        self.tape = tape.contiguous()
        self.tape_pos = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self._buf = None

    def seed_mt19937(self, seeds):
        """Per-env ``rlcard.make(..., {'seed': s})`` seeding (utils/seeding.py:33-41): the sha512 word split runs on
        the host (one digest per env), MT19937 init_by_array on the device (rlc_seed_mt19937)."""
        assert len(seeds) == self.num_envs
        keys = np.zeros((self.num_envs, 2), np.uint32)
        lens = np.zeros(self.num_envs, np.int32)
        for i, s in enumerate(seeds):
            w = seed_words(s)
            keys[i, :len(w)] = w
            lens[i] = len(w)
        dk = torch.as_tensor(keys.view(np.int32)).to(self.device)
        dl = torch.as_tensor(lens).to(self.device)
        self.mt = torch.empty((625, self.num_envs), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.rlc_seed_mt19937(_ptr(dk), _ptr(dl), self.num_envs, _ptr(self.mt), self._stream()))
        self._buf = None

    # ------------------------------------------------------------------ plumbing
    def _buffers(self):
        if self._buf is None:
            b = RlcBuffers()
            b.state = self.state.data_ptr()
            b.chance = self.chance
            b.seed = self.seed_value % 2 ** 64
            b.env_id_base = self.env_id_base
            if self.chance == CHANCE_TAPE:
                if self.tape is None:
                    raise _lib.RlcError("replay mode needs set_tape() before reset()")
                b.tape = self.tape.data_ptr(); b.tape_stride = self.tape.shape[1]; b.tape_pos = self.tape_pos.data_ptr()
            if self.chance == CHANCE_MT19937:
                if self.mt is None:
                    raise _lib.RlcError("mt19937 mode needs seed_mt19937() before reset()")
                b.mt = self.mt.data_ptr()
            b.obs = self.obs.data_ptr()
            b.obs_dtype = DTYPE_F32 if self.obs_dtype == torch.float32 else DTYPE_U8
            b.mask = self.mask.data_ptr()
            b.cur_player = self.cur_player.data_ptr()
            b.done = self.done.data_ptr()
            b.payoffs = self.payoffs.data_ptr()
            b.terminal_obs = self.terminal_obs.data_ptr() if self.terminal_obs is not None else None
            b.err = self.err.data_ptr()
            if self.legal_order is not None:
                b.legal_order, b.legal_order_stride = self.legal_order.data_ptr(), self.legal_order.shape[1]
            self._buf = b
        return self._buf

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ Env surface, vector form
    def reset(self, reset_mask=None):
        """Env.reset for all envs (or those with reset_mask != 0) -> (obs, mask, cur_player)."""
        with torch.cuda.device(self.device):
            if reset_mask is not None:
                reset_mask = reset_mask.to(device=self.device, dtype=torch.uint8).contiguous()
            check(self.L.rlc_reset(self.gid, C.byref(self._buffers()), self.num_envs, _ptr(reset_mask), self._stream()))
        self.launches += 1
        return self.obs, self.mask, self.cur_player

    def step(self, actions, auto_reset=None):
        """Env.step for all envs; actions int32 [N] on the device (negative = leave env untouched).
        -> (obs, mask, cur_player, done, payoffs)."""
        auto = self.auto_reset if auto_reset is None else auto_reset
        flags = (AUTO_RESET if auto else 0) | (TERMINAL_OBS if self.terminal_obs is not None else 0)
        with torch.cuda.device(self.device):
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
            check(self.L.rlc_step(self.gid, C.byref(self._buffers()), _ptr(actions), self.num_envs, flags, self._stream()))
        self.launches += 1
        return self.obs, self.mask, self.cur_player, self.done, self.payoffs

    def get_state(self, seat=None):
        """Env.get_state(player_id): obs/mask as seen by seat (int, int32 [N] tensor or None = current)."""
        with torch.cuda.device(self.device):
            if seat is not None and not torch.is_tensor(seat):
                seat = torch.full((self.num_envs,), int(seat), dtype=torch.int32, device=self.device)
            if seat is not None:
                seat = seat.to(device=self.device, dtype=torch.int32).contiguous()
            check(self.L.rlc_observe(self.gid, C.byref(self._buffers()), _ptr(seat), self.num_envs, self._stream()))
        self.launches += 1
        return self.obs, self.mask, self.cur_player, self.done, self.payoffs

    def alloc_host_step(self):
        """Pinned host mirror of the per-step outputs: (buffer, dict of views obs / mask / cur_player / done / payoffs)."""
        offs, sizes = self._sections
        N = self.num_envs
        buf = torch.zeros(self._out.numel(), dtype=torch.uint8).pin_memory()
        sec = lambda k: buf[offs[k]:offs[k] + sizes[k]]
        views = {'obs': sec(0).view(self.obs_dtype).view(N, self.obs_stride),
                 'mask': sec(1).view(torch.int32).view(N, self.mask_words) if self.mask_bitpacked else sec(1).view(N, self.num_actions),
                 'cur_player': sec(2).view(torch.int32), 'done': sec(3),
                 'payoffs': sec(4).view(torch.float32).view(N, self.num_players)}
        if self.legal_order is not None:
            views['legal_order'] = sec(5).view(torch.int32).view(N, self.legal_order.shape[1])
        return buf, views

    def step_host(self, host_actions, host_buf, device_actions=None):
        """Env.step for a HOST agent: actions come from (pinned) host memory, the whole post-step view (obs, mask,
        cur_player, done, payoffs) lands in ``host_buf`` (alloc_host_step) with one device-to-host copy.  Synchronous."""
        if device_actions is None:
            device_actions = getattr(self, '_h_actions', None)
            if device_actions is None:
                device_actions = self._h_actions = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        device_actions.copy_(host_actions, non_blocking=True)
        self.step(device_actions)
        host_buf.copy_(self._out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def alloc_trajectory(self, T, obs=True, mask=True):
        N, dev = self.num_envs, self.device
        tr = {}
        tr['obs'] = torch.empty((T, N, self.obs_stride), dtype=self.obs_dtype, device=dev) if obs else None
        if mask:
            shape = (T, N, self.mask_words) if self.mask_bitpacked else (T, N, self.num_actions)
            tr['mask'] = torch.empty(shape, dtype=torch.int32 if self.mask_bitpacked else torch.uint8, device=dev)
        else:
            tr['mask'] = None
        tr['action'] = torch.empty((T, N), dtype=torch.int32, device=dev)
        tr['player'] = torch.empty((T, N), dtype=torch.int32, device=dev)
        tr['done'] = torch.empty((T, N), dtype=torch.uint8, device=dev)
        tr['payoffs'] = torch.empty((T, N, self.num_players), dtype=torch.float32, device=dev)
        return tr

    def alloc_terminal_pool(self, rows, T, out=None):
        """Adds the per-seat terminal-state pool of the fused rollout (env.py:161-164) to a trajectory dict:
        terminal_obs [rows, P, obs_stride], terminal_mask [rows, A | mask_words], terminal_row int32 [T, N] (pool row of
        the episode that ended in that cell, -1 otherwise), terminal_count int32 [1]."""
        out = {} if out is None else out
        N, dev = self.num_envs, self.device
        out['terminal_obs'] = torch.zeros((rows, self.num_players, self.obs_stride), dtype=self.obs_dtype, device=dev)
        mshape = (rows, self.mask_words) if self.mask_bitpacked else (rows, self.num_actions)
        out['terminal_mask'] = torch.zeros(mshape, dtype=torch.int32 if self.mask_bitpacked else torch.uint8, device=dev)
        out['terminal_row'] = torch.full((T, N), -1, dtype=torch.int32, device=dev)
        out['terminal_count'] = torch.zeros(1, dtype=torch.int32, device=dev)
        return out

    def rollout_random(self, T, out=None, actions=None):
        """T env-steps per env with on-device uniform-random agents (the Env.run loop with RandomAgents),
        auto reset.  Returns trajectory tensors [T, N, ...]: obs/mask the acting player saw, action,
        player, done, payoffs.  ``actions`` (int32 [T, N], device): apply these recorded ids instead of the random
        policy (negative = the env idles in that cell) -- the action half of the replay mode.  If ``out`` carries a
        terminal pool (alloc_terminal_pool) the per-seat terminal states of every finished episode are written too."""
        if out is None:
            out = self.alloc_trajectory(T)
        tr = RlcTrajectory()
        for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
            t = out.get(k)
            if t is not None:
                assert t.is_contiguous() and t.shape[0] >= T
                setattr(tr, k, t.data_ptr())
        if actions is not None:
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
            assert actions.shape[0] >= T and actions.shape[1] == self.num_envs
            tr.forced_actions = actions.data_ptr()
        if out.get('terminal_row') is not None:
            assert out['terminal_row'].shape[0] >= T
            tr.terminal_obs, tr.terminal_mask = out['terminal_obs'].data_ptr(), out['terminal_mask'].data_ptr()
            tr.terminal_row, tr.terminal_count = out['terminal_row'].data_ptr(), out['terminal_count'].data_ptr()
            tr.terminal_capacity = out['terminal_obs'].shape[0]
        with torch.cuda.device(self.device):
            check(self.L.rlc_rollout_random(self.gid, C.byref(self._buffers()), C.byref(tr), self.num_envs, int(T),
                                            self._stream()))
        self.launches += 1
        return out

    def alloc_host_trajectory(self, T, pinned=True):
        """Pinned host buffers with the layout of alloc_trajectory (the destination of rollout_random_host)."""
        N = self.num_envs
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pinned)
        mshape = (T, N, self.mask_words) if self.mask_bitpacked else (T, N, self.num_actions)
        return {'obs': mk((T, N, self.obs_stride), self.obs_dtype),
                'mask': mk(mshape, torch.int32 if self.mask_bitpacked else torch.uint8),
                'action': mk((T, N), torch.int32), 'player': mk((T, N), torch.int32),
                'done': mk((T, N), torch.uint8), 'payoffs': mk((T, N, self.num_players), torch.float32)}

    def compact_words(self):
        """uint32 words per env-step of the compact wire format (0 = this game has none)."""
        return int(self.L.rlc_compact_words(self.gid))

    def alloc_host_compact(self, T, pinned=True):
        """Pinned host buffer for rollout_random_host(compact=True): uint32 (as int32) [T, N, words]."""
        w = self.compact_words()
        if w == 0:
            raise _lib.RlcError('%s has no compact wire format' % self.name)
        return torch.empty((T, self.num_envs, w), dtype=torch.int32, pin_memory=pinned)

    def rollout_random_host(self, T, host_out, chunk=16, host_state=None, compact=False):
        """The Env.run loop with random agents for a HOST consumer (what examples/run_random.py does: the
        trajectories end up in host memory).  T env-steps per env are produced in chunks of ``chunk`` steps by
        rlc_rollout_random into two device staging buffers; while chunk c+1 is being simulated, chunk c is
        copied device->host into the pinned ``host_out`` tensors on a second stream.  If ``host_state`` (a pinned
        int32 tensor shaped like ``self.state``) is given, the packed env state is uploaded from it before the
        first chunk and written back after the last one, i.e. the caller owns the state in host memory.
        ``compact=True`` (Leduc, Limit Hold'em, UNO, DouDizhu, Scout): ``host_out`` is the tensor of alloc_host_compact; every
        chunk is re-encoded on the device by rlc_compact_trajectory and only the 4 / 12 / 28 / 132 / 80-byte records cross PCIe
        (rlcard_b200.compact.expand rebuilds the dense rows on the host, bit-exact).
        Returns after everything has landed in host memory."""
        dev = self.device
        chunk = max(1, min(int(chunk), int(T)))
        if getattr(self, '_stage', None) is None or self._stage[0]['action'].shape[0] != chunk:
            self._stage = [self.alloc_trajectory(chunk), self.alloc_trajectory(chunk)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_full = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_packed = None
        if compact and self._stage_packed is None:
            w = self.compact_words()
            if w == 0:
                raise _lib.RlcError('%s has no compact wire format' % self.name)
            self._stage_packed = [torch.empty((chunk, self.num_envs, w), dtype=torch.int32, device=dev) for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        keys = ('obs', 'mask', 'action', 'player', 'done', 'payoffs')
        with torch.cuda.device(dev):
            if host_state is not None:
                self.state.copy_(host_state, non_blocking=True)
            t0, c = 0, 0
            while t0 < T:
                tc = min(chunk, T - t0)
                b = c & 1
                if c >= 2:
                    main.wait_event(self._stage_free[b])            # the copy of chunk c-2 has left this buffer
                self.rollout_random(tc, out=self._stage[b])
                if compact:
                    tr = RlcTrajectory()
                    for k in keys:
                        setattr(tr, k, self._stage[b][k].data_ptr())
                    check(self.L.rlc_compact_trajectory(self.gid, C.byref(tr), DTYPE_F32 if self.obs_dtype == torch.float32 else DTYPE_U8,
                                                        tc, self.num_envs, _ptr(self._stage_packed[b]), self._stream()))
                    self.launches += 1
                self._stage_full[b].record(main)
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(self._stage_full[b])
                    if compact:
                        host_out[t0:t0 + tc].copy_(self._stage_packed[b][:tc], non_blocking=True)
                    else:
                        for k in keys:
                            host_out[k][t0:t0 + tc].copy_(self._stage[b][k][:tc], non_blocking=True)
                    self._stage_free[b].record(self._copy_stream)
                t0 += tc
                c += 1
            if host_state is not None:
                host_state.copy_(self.state, non_blocking=True)
            self._copy_stream.synchronize()
            main.synchronize()
        return host_out

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything needed to resume this VecEnv bit-exactly: the packed env state (it carries the episode / step
        counters the Philox streams are keyed on), the replay cursors or MT19937 generators, and the error flags."""
        d = {'env_id': self.name, 'num_envs': self.num_envs, 'seed': self.seed_value, 'env_id_base': self.env_id_base,
             'chance': self.chance, 'state': self.state.clone(), 'err': self.err.clone()}
        if self.tape is not None:
            d['tape'], d['tape_pos'] = self.tape.clone(), self.tape_pos.clone()
        if self.mt is not None:
            d['mt'] = self.mt.clone()
        return d

    def load_state_dict(self, d):
        if d['env_id'] != self.name or d['num_envs'] != self.num_envs or d['chance'] != self.chance:
            raise ValueError('checkpoint is for %s x %d (chance %d)' % (d['env_id'], d['num_envs'], d['chance']))
        self.seed_value, self.env_id_base = int(d['seed']), int(d['env_id_base'])
        self.state.copy_(d['state']); self.err.copy_(d['err'])
        if 'tape' in d:
            self.tape = d['tape'].to(self.device).clone(); self.tape_pos = d['tape_pos'].to(self.device).clone()
        if 'mt' in d:
            self.mt = d['mt'].to(self.device).clone()
        self._buf = None
        self.get_state(None)                                  # refresh obs / mask / cur_player from the restored state

    # ------------------------------------------------------------------ rollout helpers (callers of the path)
    def run(self, policy, num_steps, auto_reset=True):
        """Vector form of Env.run: policy(obs, mask, cur_player) -> int32 actions [N] on the device.
        Returns (sum of payoffs per seat over finished episodes, episodes finished)."""
        obs, mask, cur = self.reset()
        total = torch.zeros(self.num_players, dtype=torch.float64, device=self.device)
        episodes = torch.zeros((), dtype=torch.int64, device=self.device)
        for _ in range(num_steps):
            actions = policy(obs, mask, cur)
            obs, mask, cur, done, pay = self.step(actions, auto_reset=auto_reset)
            d = done.bool()
            total += (pay * d.unsqueeze(1)).sum(0).double()
            episodes += d.sum()
        return total, episodes

    def check_errors(self):
        e = int(self.err.max().item()) if self.num_envs else 0
        if e:
            bad = int((self.err != 0).sum().item())
            raise _lib.RlcError('device error flags set on %d envs (OR=%d: 1 tape exhausted, 2 tape value out of '
                                'range, 4 illegal action replaced by fallback, 8 terminal-state pool full)' % (bad, e))


def random_policy(generator=None):
    """Uniform-random legal action from a dense mask with torch ops (a stand-in for learned policies;
    the fused path is VecEnv.rollout_random)."""
    def policy(obs, mask, cur_player):
        w = mask.float()
        return torch.multinomial(w, 1, generator=generator).squeeze(1).int()
    return policy


from .utils import vec_tournament  # noqa: E402,F401  (kept importable from here)
