"""CPU ORACLE -- test infrastructure, NOT product code.

ctypes binding of ``oracle/liboracle.so`` (plain-C restatement of the reference
engines, see ``oracle/orc.h``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this package;
``rlcard_b200`` never does.
"""
import ctypes as C
import hashlib
import os
import struct
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GAME_IDS = {'blackjack': 0, 'leduc-holdem': 1, 'limit-holdem': 2, 'uno': 3, 'doudizhu': 4, 'scout': 5, 'no-limit-holdem': 6}


def build(force=False):
    """Compile liboracle.so with gcc (make -C oracle)."""
    so = os.path.join(_DIR, 'liboracle.so')
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith(('.c', '.h'))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _DIR, '-s', '-B'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, i64, u8p, f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.orc_env_create.restype = vp; L.orc_env_create.argtypes = [i32]
        L.orc_env_destroy.argtypes = [vp]
        L.orc_env_set_tape.argtypes = [vp, vp, i64]
        L.orc_env_set_philox.argtypes = [vp, C.c_uint64, C.c_uint32]
        L.orc_env_set_mt.argtypes = [vp, vp, i32]
        L.orc_env_record.argtypes = [vp, vp, i64]
        L.orc_env_recorded.restype = i64; L.orc_env_recorded.argtypes = [vp]
        L.orc_env_tape_pos.restype = i64; L.orc_env_tape_pos.argtypes = [vp]
        L.orc_env_tape_err.argtypes = [vp]
        for f in ('reset', 'is_over', 'player'):
            getattr(L, 'orc_env_' + f).argtypes = [vp]
        L.orc_env_step.argtypes = [vp, i32]
        L.orc_env_legal.argtypes = [vp, vp]
        L.orc_env_obs.argtypes = [vp, i32, vp]
        L.orc_env_payoffs.argtypes = [vp, vp]
        L.orc_info.argtypes = [i32, vp, vp, vp]
        L.orc_envs_create.restype = vp; L.orc_envs_create.argtypes = [i32, i32, C.c_uint64, C.c_uint32]
        L.orc_envs_destroy.argtypes = [vp]
        L.orc_envs_rollout.restype = i64
        L.orc_envs_rollout.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, i32]
        L.orc_envs_episodes.restype = i64; L.orc_envs_episodes.argtypes = [vp]
        L.orc_holdem_strength7.restype = C.c_uint32; L.orc_holdem_strength7.argtypes = [vp]
        L.orc_philox_word.restype = C.c_uint32; L.orc_philox_word.argtypes = [C.c_uint32] * 7
        L.orc_philox4x32_10.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
        L.orc_doudizhu_set_table.argtypes = [vp, vp, vp, i32, i32, i32]
        L.orc_holdem_winners.argtypes = [vp, i32, vp]
        L.orc_doudizhu_legal_for.argtypes = [vp, i32, vp]
        L.orc_leduc_judge.argtypes = [i32] * 7 + [vp]
        L.orc_uno_encode.argtypes = [vp, i32, i32, vp]
        _LIB = L
    return _LIB


_DDZ = []


def _need_tables(gid):
    """DouDizhu: install the action table (id -> rank counts, type, weight) in the C oracle.  The table data
    (the reference's action-id contract) is shared with the product; the legal-set logic is not."""
    if gid == GAME_IDS['doudizhu'] and not _DDZ:
        from rlcard_b200 import doudizhu_table as T
        tab = T.load()
        c = np.ascontiguousarray(tab['counts'], np.uint64); t = np.ascontiguousarray(tab['type']); w = np.ascontiguousarray(tab['weight'])
        lib().orc_doudizhu_set_table(c.ctypes.data, t.ctypes.data, w.ctypes.data, len(c), T.T_BOMB, T.T_ROCKET)
        _DDZ.append((c, t, w))


def seed_words(seed):
    """rlcard/utils/seeding.py:33-113: sha512(str(seed))[:8] -> little-endian uint32 words for
    RandomState.seed(list) (leading zero words dropped, 0 -> [0])."""
    seed = seed % 2 ** 64
    h = hashlib.sha512(str(seed).encode('utf8')).digest()[:8]
    big = sum(v << (32 * i) for i, v in enumerate(struct.unpack('2I', h)))
    if big == 0:
        return [0]
    out = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        out.append(mod)
    return out


def need_tables(game):
    _need_tables(GAME_IDS[game] if isinstance(game, str) else game)


def info(game):
    gid = GAME_IDS[game] if isinstance(game, str) else game
    a, b = C.c_int(), C.c_int()
    od = (C.c_int * 4)()
    assert lib().orc_info(gid, C.byref(a), C.byref(b), od) == 0
    return a.value, b.value, list(od)[:a.value]


class OracleEnv:
    """One oracle env with the reference's Env surface (ids, numpy obs)."""

    def __init__(self, game):
        self.game = game
        self.gid = GAME_IDS[game]
        self.L = lib()
        _need_tables(self.gid)
        self.h = self.L.orc_env_create(self.gid)
        assert self.h
        self.num_players, self.num_actions, self.obs_dim = info(self.gid)
        self._keep = []

    def __del__(self):
        try:
            self.L.orc_env_destroy(self.h)
        except Exception:
            pass

    def set_tape(self, tape):
        tape = np.ascontiguousarray(tape, np.uint8)
        self._keep = [tape]
        self.L.orc_env_set_tape(self.h, tape.ctypes.data, tape.size)

    def set_philox(self, seed, env_id):
        self.L.orc_env_set_philox(self.h, seed, env_id)

    def seed(self, seed):
        """np.random.RandomState seeded exactly like rlcard.make(..., {'seed': seed})."""
        w = np.asarray(seed_words(seed), np.uint32)
        self.L.orc_env_set_mt(self.h, w.ctypes.data, w.size)

    def record(self, cap=1 << 20):
        self._rec = np.zeros(cap, np.uint8)
        self.L.orc_env_record(self.h, self._rec.ctypes.data, cap)

    def recorded(self):
        return self._rec[:self.L.orc_env_recorded(self.h)].copy()

    def tape_pos(self):
        return self.L.orc_env_tape_pos(self.h)

    def tape_err(self):
        return self.L.orc_env_tape_err(self.h)

    def reset(self):
        return self.L.orc_env_reset(self.h)

    def step(self, a):
        return self.L.orc_env_step(self.h, int(a))

    def legal_mask(self):
        m = np.zeros(self.num_actions, np.uint8)
        self.L.orc_env_legal(self.h, m.ctypes.data)
        return m

    def obs(self, seat=-1):
        o = np.zeros(1024, np.float32)
        d = self.L.orc_env_obs(self.h, seat, o.ctypes.data)
        return o[:d]

    def is_over(self):
        return bool(self.L.orc_env_is_over(self.h))

    def player(self):
        return self.L.orc_env_player(self.h)

    def payoffs(self):
        p = np.zeros(4, np.float64)
        self.L.orc_env_payoffs(self.h, p.ctypes.data)
        return p[:self.num_players]


class OracleVec:
    """Throughput-mode twin of VecEnv.rollout_random (Philox chance + policy, auto reset)."""

    def __init__(self, game, n, seed, env0=0):
        self.gid = GAME_IDS[game]
        self.L = lib()
        _need_tables(self.gid)
        self.n = n
        self.num_players, self.num_actions, od = info(self.gid)
        self.obs_stride = max(od)
        self.h = self.L.orc_envs_create(self.gid, n, seed, env0)

    def __del__(self):
        try:
            self.L.orc_envs_destroy(self.h)
        except Exception:
            pass

    def rollout(self, T, want_obs=True, nthreads=1):
        n, P, A = self.n, self.num_players, self.num_actions
        out = dict(
            obs=np.zeros((T, n, self.obs_stride), np.float32) if want_obs else None,
            mask=np.zeros((T, n, A), np.uint8) if want_obs else None,
            action=np.zeros((T, n), np.int32), player=np.zeros((T, n), np.int32),
            done=np.zeros((T, n), np.uint8), payoffs=np.zeros((T, n, P), np.float32))
        ptr = lambda a: a.ctypes.data if a is not None else None
        self.L.orc_envs_rollout(self.h, T, ptr(out['obs']), self.obs_stride, ptr(out['mask']), ptr(out['action']),
                                ptr(out['player']), ptr(out['done']), ptr(out['payoffs']), nthreads)
        return out

    def episodes(self):
        return self.L.orc_envs_episodes(self.h)
