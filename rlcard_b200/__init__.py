"""rlcard_b200 -- B200-native batched simulator for RLCard's card-game environments.

Public surface: ``make`` / ``Env`` (the reference's single-env API), ``VecEnv`` (N envs per GPU),
``vec_tournament``, ``random_policy``.  The compute path is the C-ABI CUDA library
``librlcard_b200.so`` (include/rlcard_b200.h); there is no CPU fallback.
"""
from ._lib import GAME_IDS, RlcError, SO_PATH, game_info, lib  # noqa: F401
from .env import DEFAULT_CONFIG, Env, RandomAgent, make  # noqa: F401
from .vec_env import VecEnv, random_policy, seed_words  # noqa: F401
from .utils import reorganize, remove_illegal, tournament, vec_tournament  # noqa: F401

__all__ = ['make', 'Env', 'VecEnv', 'RandomAgent', 'vec_tournament', 'tournament', 'reorganize', 'random_policy', 'game_info', 'lib',
           'GAME_IDS', 'RlcError']
