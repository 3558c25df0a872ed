"""The reference's own env tests (tests/envs/test_<game>_env.py, determism_util.py) restated against
``rlcard_b200.make`` -- same call sequences and assertions, the game running in the CUDA kernels."""
import random

import numpy as np
import pytest

import rlcard_b200
from rlcard_b200 import RandomAgent

pytestmark = pytest.mark.gpu

GAMES = ['blackjack', 'leduc-holdem', 'limit-holdem', 'uno', 'doudizhu', 'scout', 'no-limit-holdem']
OBS_SIZE = {'blackjack': 2, 'leduc-holdem': 36, 'limit-holdem': 72, 'uno': 240, 'doudizhu': 790, 'scout': 688, 'no-limit-holdem': 54}


def gather_observations(env, actions, num_rand_steps):      # tests/envs/determism_util.py:22-45
    def rand_iter(n):
        for _ in range(n + 1):
            random.randint(0, 1000)
            np.random.normal(size=100)
    rand_iter(num_rand_steps)
    state, _ = env.reset()
    rand_iter(num_rand_steps)
    idx, observations = 0, []
    while not env.is_over() and idx < len(actions):
        rand_iter(num_rand_steps)
        legals = list(state['legal_actions'].keys())
        state, _ = env.step(legals[actions[idx] % len(legals)])
        idx += 1
        if not env.game.is_over():
            observations.append(state)
    return observations


@pytest.mark.parametrize('game', GAMES)
def test_is_deterministic(game):                           # determism_util.py:47-60
    env = rlcard_b200.make(game)
    actions = [random.randrange(env.game.get_num_actions()) for _ in range(25)]
    hashes = []
    for rand_iters in range(2):
        env = rlcard_b200.make(game, config={'seed': 12941})
        hashes.append(hash(tuple(hash(o['obs'].tobytes()) for o in gather_observations(env, actions, rand_iters))))
    assert hashes[0] == hashes[1]


@pytest.mark.parametrize('game', GAMES)
def test_reset_and_extract_state(game):
    env = rlcard_b200.make(game)
    state, player_id = env.reset()
    assert state['obs'].size == OBS_SIZE[game] or (game == 'doudizhu' and state['obs'].size in (790, 901))
    assert player_id == env.get_player_id()
    for action in state['legal_actions']:
        assert action < env.num_actions
    if game == 'blackjack':
        assert all(score <= 30 for score in state['obs'])


@pytest.mark.parametrize('game', GAMES)
def test_get_legal_actions_and_decode(game):
    env = rlcard_b200.make(game)
    state, _ = env.reset()
    legal = env._get_legal_actions()
    assert list(legal.keys()) == list(state['legal_actions'].keys())
    for action in legal:
        assert env._decode_action(action) in env.actions
    if game == 'blackjack':
        assert env._decode_action(0) == 'hit' and env._decode_action(1) == 'stand'
    if game in ('leduc-holdem', 'limit-holdem'):            # illegal ids fall back like envs/leducholdem.py:90-96
        assert env._decode_action(3) in ('check', 'fold')


@pytest.mark.parametrize('game', GAMES)
def test_step(game):
    env = rlcard_b200.make(game)
    state, player_id = env.reset()
    assert player_id == env.get_player_id()
    action = list(state['legal_actions'].keys())[0]
    _, player_id = env.step(action)
    assert player_id == env.get_player_id()
    assert env.timestep == 1


def test_doudizhu_pass_moves_to_next_player():              # test_doudizhu_env.py test_step (a lead cannot pass -> fallback)
    env = rlcard_b200.make('doudizhu')
    state, player_id = env.reset()
    a = list(state['legal_actions'].keys())[0]
    _, nxt = env.step(a)
    assert nxt == (player_id + 1) % 3
    assert 27471 in env._get_legal_actions()                 # the follower may pass


@pytest.mark.parametrize('game', GAMES)
def test_step_back_is_off(game):
    env = rlcard_b200.make(game)
    with pytest.raises(Exception):
        env.step_back()


@pytest.mark.parametrize('game', GAMES)
@pytest.mark.parametrize('is_training', [False, True])
def test_run(game, is_training):
    env = rlcard_b200.make(game, config={'seed': 7})
    env.set_agents([RandomAgent(env.num_actions) for _ in range(env.num_players)])
    trajectories, payoffs = env.run(is_training=is_training)
    assert len(trajectories) == env.num_players
    if game in ('leduc-holdem', 'limit-holdem', 'uno', 'no-limit-holdem'):
        assert sum(payoffs) == 0
    if game == 'blackjack':
        assert payoffs[0] in (-1, 0, 1)
    if game == 'doudizhu':                                  # landlord (seat 0) wins alone or both peasants win
        assert list(payoffs) in ([1, 0, 0], [0, 1, 1])
    for pid, traj in enumerate(trajectories):               # [s0, a0, s1, ..., terminal state] per seat
        assert isinstance(traj[-1], dict) and len(traj) % 2 == 1


def test_blackjack_payoffs_over_many_episodes():            # test_blackjack_env.py test_get_payoffs
    env = rlcard_b200.make('blackjack', config={'seed': 3})
    for _ in range(40):
        env.reset()
        while not env.is_over():
            env.step(int(np.random.choice([0, 1])))
        assert env.get_payoffs()[0] in (-1, 0, 1)


def test_unsupported_variants_raise():
    with pytest.raises(NotImplementedError):
        rlcard_b200.make('blackjack', config={'game_num_players': 5})
    with pytest.raises(NotImplementedError):
        rlcard_b200.make('leduc-holdem', config={'allow_step_back': True})
    for cfg in ({'hand_size': 12}, {'rank_count': 8}, {'game_num_players': 3}):     # envs/scout.py:14-28: refused, never ignored
        with pytest.raises(NotImplementedError):
            rlcard_b200.make('scout', config=cfg)
    with pytest.raises(NotImplementedError):
        rlcard_b200.make('limit-holdem', config={'game_num_players': 3})
    with pytest.raises(ValueError):
        rlcard_b200.make('uno', config={'no_such_key': 1})
    rlcard_b200.make('scout', config={'hand_size': 16, 'rank_count': 10, 'game_num_players': 4})      # the defaults pass


@pytest.mark.parametrize('game', GAMES)
def test_get_perfect_information(game):
    env = rlcard_b200.make(game, config={'seed': 11})
    _, player_id = env.reset()
    info = env.get_perfect_information()
    assert info['current_player'] == player_id
    assert info['legal_actions'] == env.game.get_legal_actions()
    if game == 'limit-holdem':
        assert info['chips'] in ([1, 2], [2, 1]) and info['public_card'] is None and len(info['hand_cards'][0]) == 2
    if game == 'leduc-holdem':
        assert info['chips'] in ([1, 2], [2, 1]) and info['public_card'] is None and info['current_round'] == 0


@pytest.mark.gpu
@pytest.mark.parametrize('game', ['uno', 'scout', 'leduc-holdem', 'limit-holdem', 'blackjack', 'no-limit-holdem'])
def test_random_agent_draws_replay_exactly(game):
    """examples/run_random.py with the facade: the legal ids come in the reference's insertion order (UNO hand order,
    Scout enumeration order), so RandomAgent's np.random.choice picks the same actions as on the reference under the
    same env seed and global numpy seed (tests/golden/legal_order.npz, recorded from the live reference)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'legal_order.npz'))
    k = game.replace('-', '_')
    ids, off, acts, pays = z[k + '_ids'], z[k + '_off'], z[k + '_actions'], z[k + '_payoffs']
    env = rlcard_b200.make(game, {'seed': int(z['env_seed'])})
    seen = []

    class Agent(rlcard_b200.RandomAgent):
        use_raw = False

        def step(self, state):
            i = len(seen)
            assert list(state['legal_actions'].keys()) == ids[off[i]:off[i + 1]].tolist(), (game, 'decision', i)
            a = rlcard_b200.RandomAgent.step(state)
            seen.append(a)
            return a
    env.set_agents([Agent(env.num_actions) for _ in range(env.num_players)])
    np.random.seed(int(z['np_seed']))
    for ep in range(len(pays)):
        _, payoffs = env.run(is_training=True)
        np.testing.assert_array_equal(np.asarray(payoffs, np.float64), pays[ep], err_msg='%s episode %d' % (game, ep))
    assert seen == acts.tolist()
