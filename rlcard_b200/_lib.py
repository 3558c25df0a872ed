"""ctypes binding of the C-ABI CUDA library (include/rlcard_b200.h).

There is no CPU fallback: if ``librlcard_b200.so`` is missing or fails to load, importing
any compute entry point raises.  Build it with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C rlcard_b200/csrc -j``.
"""
import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
# RLC_SO_VARIANT=tma loads the measurement build whose tile flushes are TMA bulk copies (csrc/Makefile TMA=1)
_VARIANT = os.environ.get('RLC_SO_VARIANT', '')
SO_PATH = os.path.join(_DIR, 'librlcard_b200%s.so' % ('_' + _VARIANT if _VARIANT else ''))

RLC_MAX_PLAYERS = 4
GAME_IDS = {'blackjack': 0, 'leduc-holdem': 1, 'limit-holdem': 2, 'uno': 3, 'doudizhu': 4, 'scout': 5, 'no-limit-holdem': 6}
CHANCE_PHILOX, CHANCE_TAPE, CHANCE_MT19937 = 0, 1, 2
DTYPE_U8, DTYPE_F32 = 0, 1
AUTO_RESET, TERMINAL_OBS = 1, 2
STATE_SOA, STATE_ROWS = 0, 1
ABI_VERSION = 2


class RlcInfo(C.Structure):
    _fields_ = [('game_id', C.c_int32), ('num_players', C.c_int32), ('num_actions', C.c_int32),
                ('obs_dim', C.c_int32 * RLC_MAX_PLAYERS), ('obs_stride', C.c_int32),
                ('obs_native_dtype', C.c_int32), ('mask_bitpacked', C.c_int32), ('mask_words', C.c_int32),
                ('state_words', C.c_int32), ('max_tape_draws_reset', C.c_int32), ('threads_per_env', C.c_int32),
                ('state_layout', C.c_int32), ('state_words_philox', C.c_int32), ('reserved', C.c_int32 * 2)]


class RlcBuffers(C.Structure):
    _fields_ = [('state', C.c_void_p), ('chance', C.c_int32), ('seed', C.c_uint64), ('env_id_base', C.c_uint32),
                ('tape', C.c_void_p), ('tape_stride', C.c_int32), ('tape_pos', C.c_void_p), ('mt', C.c_void_p),
                ('obs', C.c_void_p), ('obs_dtype', C.c_int32), ('mask', C.c_void_p), ('cur_player', C.c_void_p),
                ('done', C.c_void_p), ('payoffs', C.c_void_p), ('terminal_obs', C.c_void_p), ('err', C.c_void_p),
                ('legal_order', C.c_void_p), ('legal_order_stride', C.c_int32)]


class RlcTrajectory(C.Structure):
    _fields_ = [('obs', C.c_void_p), ('mask', C.c_void_p), ('action', C.c_void_p), ('player', C.c_void_p),
                ('done', C.c_void_p), ('payoffs', C.c_void_p),
                ('forced_actions', C.c_void_p), ('terminal_obs', C.c_void_p), ('terminal_mask', C.c_void_p),
                ('terminal_row', C.c_void_p), ('terminal_count', C.c_void_p), ('terminal_capacity', C.c_int32)]


class RlcDmcBuffers(C.Structure):
    _fields_ = [('open_obs', C.c_void_p), ('open_action', C.c_void_p), ('open_player', C.c_void_p), ('open_len', C.c_void_p),
                ('open_capacity', C.c_int32),
                ('out_state', C.c_void_p * RLC_MAX_PLAYERS), ('out_action', C.c_void_p * RLC_MAX_PLAYERS),
                ('out_target', C.c_void_p * RLC_MAX_PLAYERS), ('out_episode_return', C.c_void_p * RLC_MAX_PLAYERS),
                ('out_done', C.c_void_p * RLC_MAX_PLAYERS), ('out_count', C.c_void_p), ('out_capacity', C.c_int32),
                ('overflow', C.c_void_p)]


class RlcRlBuffers(C.Structure):
    _fields_ = [('pend_obs', C.c_void_p), ('pend_action', C.c_void_p), ('pend_valid', C.c_void_p),
                ('out_state', C.c_void_p * RLC_MAX_PLAYERS), ('out_action', C.c_void_p * RLC_MAX_PLAYERS),
                ('out_reward', C.c_void_p * RLC_MAX_PLAYERS), ('out_next_state', C.c_void_p * RLC_MAX_PLAYERS),
                ('out_next_mask', C.c_void_p * RLC_MAX_PLAYERS), ('out_done', C.c_void_p * RLC_MAX_PLAYERS),
                ('out_count', C.c_void_p), ('out_capacity', C.c_int32), ('overflow', C.c_void_p)]


EXPORTS = ['rlc_abi_version', 'rlc_last_error', 'rlc_game_info', 'rlc_upload_tables', 'rlc_reset', 'rlc_step',
           'rlc_observe', 'rlc_rollout_random', 'rlc_launch_count', 'rlc_judge_holdem', 'rlc_judge_leduc',
           'rlc_judge_doudizhu', 'rlc_encode_uno', 'rlc_dmc_collect', 'rlc_rl_feed', 'rlc_legal_ids', 'rlc_action_features',
           'rlc_reorganize', 'rlc_seed_mt19937', 'rlc_compact_words', 'rlc_compact_trajectory']

_LIB = None


class RlcError(RuntimeError):
    pass


def lib():
    """Load the CUDA extension; raises (no fallback) when it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RlcError('CUDA extension %s is not built (run __graft_entry__.build()); there is no CPU fallback'
                           % SO_PATH)
        L = C.CDLL(SO_PATH)
        vp, i32 = C.c_void_p, C.c_int
        L.rlc_abi_version.restype = i32
        L.rlc_last_error.restype = C.c_char_p
        L.rlc_launch_count.restype = C.c_int64
        L.rlc_game_info.argtypes = [i32, C.POINTER(RlcInfo)]
        L.rlc_upload_tables.argtypes = [i32, i32, vp, C.c_size_t]
        L.rlc_reset.argtypes = [i32, C.POINTER(RlcBuffers), i32, vp, vp]
        L.rlc_step.argtypes = [i32, C.POINTER(RlcBuffers), vp, i32, i32, vp]
        L.rlc_observe.argtypes = [i32, C.POINTER(RlcBuffers), vp, i32, vp]
        L.rlc_rollout_random.argtypes = [i32, C.POINTER(RlcBuffers), C.POINTER(RlcTrajectory), i32, i32, vp]
        L.rlc_judge_holdem.argtypes = [vp, i32, i32, vp, vp]
        L.rlc_judge_leduc.argtypes = [vp, i32, vp, vp]
        L.rlc_judge_doudizhu.argtypes = [vp, vp, i32, vp, vp]
        L.rlc_encode_uno.argtypes = [vp, vp, i32, vp, vp]
        L.rlc_legal_ids.argtypes = [i32, vp, i32, i32, vp, vp, vp]
        L.rlc_action_features.argtypes = [i32, vp, i32, vp, vp]
        L.rlc_rl_feed.argtypes = [i32, i32, C.POINTER(RlcBuffers), vp, i32, C.POINTER(RlcRlBuffers), vp]
        L.rlc_dmc_collect.argtypes = [i32, C.POINTER(RlcTrajectory), i32, i32, i32, C.POINTER(RlcDmcBuffers), vp]
        L.rlc_reorganize.argtypes = [i32, C.POINTER(RlcTrajectory), i32, i32, i32, C.POINTER(RlcRlBuffers), vp]
        L.rlc_seed_mt19937.argtypes = [vp, vp, i32, vp, vp]
        L.rlc_compact_words.argtypes = [i32]
        L.rlc_compact_trajectory.argtypes = [i32, C.POINTER(RlcTrajectory), i32, i32, i32, vp, vp]
        if L.rlc_abi_version() != ABI_VERSION:
            raise RlcError('%s has ABI %d, the Python side expects %d: rebuild (__graft_entry__.build())'
                           % (SO_PATH, L.rlc_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise RlcError('rlcard_b200 C-ABI call failed (%d): %s' % (rc, lib().rlc_last_error().decode()))


def game_info(game):
    gid = GAME_IDS[game] if isinstance(game, str) else int(game)
    info = RlcInfo()
    check(lib().rlc_game_info(gid, C.byref(info)))
    return info
