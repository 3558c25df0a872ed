"""Shared replay harness: drives an env adapter through a golden fixture
(tests/golden/<game>.npz, recorded from the live reference) and checks every
record bit-exactly.  Used for the CPU oracle and, through the C-ABI, the CUDA path."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ALL_GAMES = ['blackjack', 'leduc-holdem', 'limit-holdem', 'uno', 'doudizhu', 'scout', 'no-limit-holdem']


# extra fixtures of a game that pin rare paths: name -> env id
EXTRA_FIXTURES = {'uno-reshuffle': 'uno'}
ALL_FIXTURES = ALL_GAMES + list(EXTRA_FIXTURES)


def fixture_game(name):
    return EXTRA_FIXTURES.get(name, name)


def have_fixture(game):
    return os.path.exists(os.path.join(GOLDEN, game.replace('-', '_') + '.npz'))


def load_fixture(game):
    z = np.load(os.path.join(GOLDEN, game.replace('-', '_') + '.npz'))
    fx = {k: z[k] for k in z.files}
    fx['num_players'] = int(fx['num_players']); fx['num_actions'] = int(fx['num_actions'])
    A = fx['num_actions']
    fx['legal'] = np.unpackbits(fx['rec_legal'], axis=1, bitorder='little')[:, :A]
    return fx


def slot_records(fx, slot):
    return np.nonzero(fx['rec_slot'] == slot)[0]


def slot_tape(fx, slot):
    return fx['tape'][fx['tape_off'][slot]:fx['tape_off'][slot + 1]]


def check_slot(fx, slot, env, what=''):
    """env: adapter with reset() step(a) legal_mask() obs(seat) is_over() player() payoffs();
    obs(-1) = obs of the state returned by the last reset/step."""
    n = 0
    for r in slot_records(fx, slot):
        kind, arg = int(fx['rec_kind'][r]), int(fx['rec_arg'][r])
        tag = '%s slot %d record %d kind %d arg %d' % (what, slot, r, kind, arg)
        if kind == 0:
            env.reset()
        elif kind == 1:
            env.step(arg)
        if kind == 3:
            np.testing.assert_array_equal(np.asarray(env.payoffs(), np.float64), fx['rec_payoffs'][r], err_msg=tag)
        else:
            d = int(fx['rec_obs_dim'][r])
            got = np.asarray(env.obs(arg if kind == 2 else -1))
            assert got.shape[0] == d, tag
            np.testing.assert_array_equal(got.astype(np.float64), fx['rec_obs'][r, :d].astype(np.float64), err_msg=tag)
            np.testing.assert_array_equal(np.asarray(env.legal_mask()), fx['legal'][r], err_msg=tag)
        assert env.player() == fx['rec_player'][r], tag
        assert bool(env.is_over()) == bool(fx['rec_done'][r]), tag
        n += 1
    return n
