"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, reports the reference's env metadata, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

import rlcard_b200
from rlcard_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'rlcard_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rlc_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.SO_PATH)
    names = header_functions()
    assert set(names) >= set(_lib.EXPORTS)
    for name in names:
        assert hasattr(L, name), name


def test_abi_version_and_info():
    L = rlcard_b200.lib()
    assert L.rlc_abi_version() == 2
    # Env.num_players / num_actions / state_shape of the reference (envs/*.py)
    expect = {'blackjack': (1, 2, [2]), 'leduc-holdem': (2, 4, [36, 36]), 'limit-holdem': (2, 4, [72, 72]),
              'uno': (2, 61, [240, 240]), 'doudizhu': (3, 27472, [790, 901, 901]), 'scout': (4, 204, [688] * 4),
              'no-limit-holdem': (2, 5, [54, 54])}
    for g, (p, a, od) in expect.items():
        i = rlcard_b200.game_info(g)
        assert (i.num_players, i.num_actions, list(i.obs_dim)[:p]) == (p, a, od)
    bad = _lib.RlcInfo()
    assert L.rlc_game_info(99, ctypes.byref(bad)) < 0
    assert b'bad game id' in L.rlc_last_error()


def test_argument_errors_do_not_launch():
    L = rlcard_b200.lib()
    b = _lib.RlcBuffers()
    assert L.rlc_reset(1, ctypes.byref(b), 4, None, None) == -1      # null state
    assert L.rlc_step(1, ctypes.byref(b), None, 0, 0, None) == -1
    assert L.rlc_launch_count() == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(Exception):
        rlcard_b200.VecEnv('leduc-holdem', 4, device='cpu')
    with pytest.raises(Exception):
        rlcard_b200.VecEnv('leduc-holdem', 4, device='cuda:0').reset()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'rlcard_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'liboracle' not in txt and 'orc.h' not in txt, f


def test_reorganize_matches_reference_outputs():
    """rlcard/utils/utils.py:153-179 on fixtures produced by the live reference (tests/golden/make_reorganize_golden.py)."""
    import json
    from rlcard_b200.utils import reorganize
    cases = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reorganize.json')))
    assert len(cases) >= 20
    for c in cases:
        assert reorganize(c['trajectories'], c['payoffs']) == c['expected']


def test_tournament_averages_over_games():
    from rlcard_b200.utils import tournament

    class FakeEnv:
        num_players = 2
        calls = 0

        def run(self, is_training=False):
            FakeEnv.calls += 1
            import numpy as np
            return None, np.array([1.0, -1.0]) if FakeEnv.calls % 2 else [np.array([0.0, 0.0]), np.array([2.0, -2.0])]
    assert tournament(FakeEnv(), 6) == [1.0, -1.0]
