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


# ---- deep fixtures (tests/golden/deep_<game>.npz, make_deep_golden.py): obs rows / legal sets as 32-bit digests ----
import hashlib  # noqa: E402

DEEP_GAMES = [g for g in ALL_GAMES if os.path.exists(os.path.join(GOLDEN, 'deep_' + g.replace('-', '_') + '.npz'))]


def _h32(b):
    return int.from_bytes(hashlib.blake2b(b, digest_size=4).digest(), 'little')


def obs_digest(obs_row):
    return _h32(np.ascontiguousarray(obs_row, dtype='<f4').tobytes())


def legal_digest(mask_row):
    return _h32(np.flatnonzero(mask_row).astype('<u2').tobytes())


def load_deep(game):
    z = np.load(os.path.join(GOLDEN, 'deep_' + game.replace('-', '_') + '.npz'))
    fx = {k: z[k] for k in z.files}
    fx['num_players'] = int(fx['num_players']); fx['num_actions'] = int(fx['num_actions'])
    pay = np.zeros((len(fx['rec_kind']), fx['num_players']), np.float64)
    pay[fx['pay_rec']] = fx['pay']
    fx['rec_payoffs'] = pay
    return fx


def deep_slot_records(fx, slot):
    return range(int(fx['rec_off'][slot]), int(fx['rec_off'][slot + 1]))


def check_record_deep(fx, r, obs_row, mask_row, tag):
    """obs_row: at least rec_obs_dim values (anything beyond must be zero padding); mask_row: dense 0/1 [A]."""
    d = int(fx['rec_obs_dim'][r])
    assert not np.any(obs_row[d:]), tag + ' (padding not zero)'
    assert int(np.count_nonzero(mask_row)) == int(fx['rec_nlegal'][r]), tag + ' (legal count)'
    assert legal_digest(mask_row) == int(fx['rec_legal_hash'][r]), tag + ' (legal set)'
    assert obs_digest(obs_row[:d]) == int(fx['rec_obs_hash'][r]), tag + ' (obs)'


def check_slot_deep(fx, slot, env, what=''):
    n = 0
    for r in deep_slot_records(fx, slot):
        kind, arg = int(fx['rec_kind'][r]), int(fx['rec_arg'][r])
        tag = '%s slot %d record %d kind %d arg %d' % (what, slot, r, kind, arg)
        if kind == 0:
            env.reset()
        elif kind == 1:
            env.step(arg)
        if kind == 3:
            np.testing.assert_array_equal(np.asarray(env.payoffs(), np.float64), fx['rec_payoffs'][r], err_msg=tag)
        else:
            check_record_deep(fx, r, np.asarray(env.obs(arg if kind == 2 else -1)), np.asarray(env.legal_mask()), tag)
        assert env.player() == fx['rec_player'][r], tag
        assert bool(env.is_over()) == bool(fx['rec_done'][r]), tag
        n += 1
    return n
