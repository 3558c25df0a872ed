"""Single-env facade with the reference's ``Env`` surface (rlcard/envs/env.py:9-231).

``rlcard_b200.make(env_id, config)`` returns an object whose ``reset/step/get_state/is_over/
get_player_id/get_payoffs/run/set_agents/seed/get_action_feature`` behave like the reference's, with the
game itself running in the CUDA kernels (a ``VecEnv`` of one env in ``mt19937`` mode: the same
``np.random.RandomState`` stream the reference would consume for this seed, so trajectories are
identical to ``rlcard.make(env_id, {'seed': s})``).
"""
from collections import OrderedDict

import numpy as np
import torch

from .vec_env import VecEnv

DEFAULT_CONFIG = {'allow_step_back': False, 'seed': None}   # envs/registration.py:4-7

# per env: raw action names by id, reference obs dtype, reference obs shape per seat
_POKER_ACTIONS = ['call', 'raise', 'fold', 'check']


def _uno_actions():
    colors, traits = ['r', 'g', 'b', 'y'], [str(i) for i in range(10)] + ['skip', 'reverse', 'draw_2', 'wild', 'wild_draw_4']
    return ['%s-%s' % (c, t) for c in colors for t in traits] + ['draw']   # games/uno/jsondata/action_space.json


def _scout_actions():
    acts = []
    for s in range(16):
        for e in range(s + 1, 17):
            acts.append('play-%d-%d' % (s, e))                             # games/scout/utils/utils.py:186-200
    for ins in range(17):
        for tag in ('front-%d-normal', 'front-%d-flip', 'back-%d-normal', 'back-%d-flip'):
            acts.append('scout-' + tag % ins)
    return acts


def _doudizhu_actions():
    from . import doudizhu_table
    return doudizhu_table.action_strings()                                 # games/doudizhu/utils.py:21-27 ID_2_ACTION


_SPECS = {
    'blackjack': dict(actions=lambda: ['hit', 'stand'], dtype=np.int64, shape=lambda d: (d,)),
    'leduc-holdem': dict(actions=lambda: _POKER_ACTIONS, dtype=np.float64, shape=lambda d: (d,)),
    'limit-holdem': dict(actions=lambda: _POKER_ACTIONS, dtype=np.float64, shape=lambda d: (d,)),
    'uno': dict(actions=_uno_actions, dtype=np.int64, shape=lambda d: (4, 4, 15)),
    'doudizhu': dict(actions=lambda: _doudizhu_actions(), dtype=np.int8, shape=lambda d: (d,)),
    'scout': dict(actions=_scout_actions, dtype=np.float32, shape=lambda d: (d,)),
    # games/nolimitholdem/round.py:8-18 Action enum names (the reference's raw actions are the enum members)
    'no-limit-holdem': dict(actions=lambda: ['FOLD', 'CHECK_CALL', 'RAISE_HALF_POT', 'RAISE_POT', 'ALL_IN'], dtype=np.float64,
                            shape=lambda d: (d,)),
}
_STATE_SHAPE = {'uno': [4, 4, 15]}
_INT_PAYOFF = {'blackjack', 'uno', 'doudizhu', 'scout', 'no-limit-holdem'}


class Env:
    """The reference's Env API over one CUDA-resident env."""

    def __init__(self, env_id, config=None, device='cuda:0'):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        # envs/env.py:30-39 forwards game_* keys to game.configure (blackjack, leduc, limit, no-limit), envs/scout.py:14-28
        # reads hand_size / rank_count.  The kernels are built for the reference's default geometry only: anything else is
        # refused loudly (never silently ignored).
        defaults = {'game_num_players': {'blackjack': 1, 'doudizhu': 3, 'scout': 4}.get(env_id, 2), 'game_num_decks': 1,
                    'hand_size': 16, 'rank_count': 10, 'chips_for_each': 100, 'dealer_id': None}
        for k, v in cfg.items():
            if k in ('allow_step_back', 'seed'):
                continue
            if k not in defaults:
                raise ValueError('unknown config key %r' % k)
            if v != defaults[k]:
                raise NotImplementedError('%s=%r: only the reference default (%r) is built' % (k, v, defaults[k]))
        self.name = env_id
        self.allow_step_back = cfg['allow_step_back']
        if self.allow_step_back:
            raise NotImplementedError('allow_step_back=True is not supported by the batched simulator')
        self.device = device
        self._spec = _SPECS[env_id]
        self._vec = VecEnv(env_id, 1, device=device, mode='mt19937', auto_reset=False, legal_order=True)
        self._hbuf, self._h = self._vec.alloc_host_step()      # pinned mirror: one D2H per reset / step / get_state
        self.num_players = self._vec.num_players
        self.num_actions = self._vec.num_actions
        self.state_shape = [list(_STATE_SHAPE.get(env_id, [d])) for d in self._vec.obs_dims]
        self.action_shape = [[54] if env_id == 'doudizhu' else None for _ in range(self.num_players)]
        self.actions = self._spec['actions']()
        if env_id == 'doudizhu':
            from . import doudizhu_table
            self._features = doudizhu_table.load()['features']
        self.timestep = 0
        self.action_recorder = []
        self.agents = None
        self._h_action = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._d_action = torch.zeros(1, dtype=torch.int32, device=self._vec.device)
        self._pid, self._done, self._pay, self._cur_state = 0, False, np.zeros(self.num_players), None
        self.seed(cfg['seed'])

    # -- seeding (env.py:228-231, utils/seeding.py:33-41)
    def seed(self, seed=None):
        if seed is None:
            import os
            seed = int.from_bytes(os.urandom(8), 'little')
        if not (isinstance(seed, int) and seed >= 0):
            raise ValueError('Seed must be a non-negative integer or omitted, not {}'.format(seed))
        self._seed = seed
        self._vec.seed_mt19937([seed])
        return seed

    # -- state dict assembly (the per-env _extract_state)
    def _fetch(self):
        """One device-to-host copy of everything the last launch wrote (obs, mask, ordered legal ids, player, done,
        payoffs); is_over / get_player_id / get_payoffs then answer from this host copy."""
        self._hbuf.copy_(self._vec._out, non_blocking=True)
        torch.cuda.current_stream(self._vec.device).synchronize()
        h = self._h
        self._pid, self._done = int(h['cur_player'][0]), bool(h['done'][0])
        self._pay = h['payoffs'][0].numpy().astype(np.float64)

    def _state_dict(self):
        self._fetch()
        self._cur_state = self._dict_from(self._pid)
        return self._cur_state, self._pid

    def _dict_from(self, seat_for_dim):
        h = self._h
        d = self._vec.obs_dims[seat_for_dim]
        obs = h['obs'][0, :d].numpy().astype(self._spec['dtype']).reshape(self._spec['shape'](d))
        order = h['legal_order'][0].numpy()
        ids = order[order >= 0].tolist()                  # the reference's insertion order (rlc_buffers.legal_order)
        if self.name == 'doudizhu':                                        # envs/doudizhu.py:112-120: id -> 54-d feature
            state = {'obs': obs, 'legal_actions': OrderedDict((int(a), self._features[a]) for a in ids)}
        else:
            state = {'obs': obs, 'legal_actions': OrderedDict((int(a), None) for a in ids)}
        state['raw_legal_actions'] = [self.actions[a] for a in ids] if self.actions else list(ids)
        state['raw_obs'] = {'obs': obs, 'legal_actions': state['raw_legal_actions']}
        state['action_record'] = self.action_recorder
        return state

    def reset(self):
        self._vec.reset()
        self.action_recorder = []
        return self._state_dict()

    def step(self, action, raw_action=False):
        if raw_action:
            action = self.actions.index(action)
        self.timestep += 1
        self.action_recorder.append((self.get_player_id(), self.actions[action] if self.actions else action))
        self._h_action[0] = int(action)
        self._d_action.copy_(self._h_action, non_blocking=True)
        self._vec.step(self._d_action, auto_reset=False)       # one launch; _state_dict makes the one D2H
        return self._state_dict()

    def step_back(self):
        raise Exception('Step back is off. To use step_back, please set allow_step_back=True in rlcard.make')

    def set_agents(self, agents):
        self.agents = agents

    def is_over(self):
        return self._done

    def get_player_id(self):
        return self._pid

    def get_state(self, player_id):
        self._vec.get_state(int(player_id))
        self._fetch()
        return self._dict_from(int(player_id))

    def get_payoffs(self):
        p = self._pay
        return p.astype(np.int64) if self.name in _INT_PAYOFF else p.copy()

    def get_action_feature(self, action):
        if self.name == 'doudizhu':
            return self._features[action].copy()               # envs/doudizhu.py:136-142
        feature = np.zeros(self.num_actions, dtype=np.int8)   # env.py:217-226 default one-hot
        feature[action] = 1
        return feature

    # -- the per-env helpers the reference's callers and tests touch (envs/<game>.py)
    def _get_legal_actions(self):
        """OrderedDict of the current player's legal action ids (envs/<game>.py _get_legal_actions)."""
        if self._cur_state is None:                      # before the first reset (blackjack's env tests do this): all ids
            return OrderedDict((a, None) for a in range(self.num_actions))
        return self._cur_state['legal_actions']

    def _decode_action(self, action_id):
        """id -> raw action with the reference's per-env fallback for illegal ids: poker 'check' else 'fold'
        (envs/leducholdem.py:81-96), UNO a uniform legal action from the global np.random (envs/uno.py:39-45)."""
        if self.name in ('leduc-holdem', 'limit-holdem'):
            legal = [self.actions[a] for a in self._get_legal_actions()]
            if self.actions[action_id] not in legal:
                return 'check' if 'check' in legal else 'fold'
        elif self.name == 'uno':
            legal = list(self._get_legal_actions().keys())
            if action_id not in legal:
                return self.actions[int(np.random.choice(legal))]
        return self.actions[action_id]

    def get_perfect_information(self):
        """envs/<game>.py get_perfect_information from the packed device state.  Exact for limit-holdem; Leduc
        cards are reported by rank only (the packed state drops the suit, which never influences play); the other
        games report the fields that do not need a decoder (current player, legal actions)."""
        pid = self.get_player_id()
        legal = [self.actions[a] for a in self._get_legal_actions()]
        info = {'current_player': pid, 'legal_actions': legal}
        st = self._vec.state[0] if self._vec.state_rows else self._vec.state[:, 0]
        w = [int(x) & 0xffffffff for x in st.cpu().tolist()][3:]                       # game words after the header
        if self.name == 'leduc-holdem':
            g = w[0]
            info['chips'] = [(g >> 7) & 15, (g >> 11) & 15]
            info['hand_cards'] = ['JQK'[g & 3], 'JQK'[(g >> 2) & 3]]
            info['public_card'] = 'JQK'[(g >> 4) & 3] if (g >> 6) & 1 else None
            info['current_round'] = (g >> 28) & 3
        elif self.name == 'limit-holdem':
            cards = [(w[0] >> (6 * k)) & 63 for k in range(5)] + [(w[1] >> (6 * k)) & 63 for k in range(4)]
            name = lambda c: 'SHDC'[c // 13] + 'A23456789TJQK'[c % 13]
            rc = (w[2] >> 22) & 7
            n_public = {0: 0, 1: 3, 2: 4}.get(rc, 5)
            info['chips'] = [w[2] & 63, (w[2] >> 6) & 63]
            info['hand_cards'] = [[name(cards[0]), name(cards[2])], [name(cards[1]), name(cards[3])]]
            info['public_card'] = [name(c) for c in cards[4:4 + n_public]] or None
        return info

    @property
    def game(self):
        """Minimal stand-in for ``env.game`` (the Python engine object does not exist: the game runs on the GPU)."""
        return _GameView(self)

    # -- the rollout loop, env.py:120-169 verbatim in behaviour
    def run(self, is_training=False):
        trajectories = [[] for _ in range(self.num_players)]
        state, player_id = self.reset()
        trajectories[player_id].append(state)
        while not self.is_over():
            if not is_training:
                action, _ = self.agents[player_id].eval_step(state)
            else:
                action = self.agents[player_id].step(state)
            next_state, next_player_id = self.step(action, self.agents[player_id].use_raw)
            trajectories[player_id].append(action)
            state, player_id = next_state, next_player_id
            if not self.is_over():
                trajectories[player_id].append(state)
        for pid in range(self.num_players):
            trajectories[pid].append(self.get_state(pid))
        return trajectories, self.get_payoffs()


class _GameView:
    """The few ``env.game`` methods the reference's rollout helpers and tests call."""

    def __init__(self, env):
        self._env = env

    def get_num_players(self):
        return self._env.num_players

    def get_num_actions(self):
        return self._env.num_actions

    def is_over(self):
        return self._env.is_over()

    def get_player_id(self):
        return self._env.get_player_id()

    def get_legal_actions(self):
        return [self._env.actions[a] for a in self._env._get_legal_actions()]

    def get_payoffs(self):
        return self._env.get_payoffs()


class RandomAgent:
    """agents/random_agent.py:14-47: uniform over state['legal_actions'] with the global np.random."""
    use_raw = False

    def __init__(self, num_actions):
        self.num_actions = num_actions

    @staticmethod
    def step(state):
        return int(np.random.choice(list(state['legal_actions'].keys())))

    def eval_step(self, state):
        n = len(state['legal_actions'])
        info = {'probs': {state['raw_legal_actions'][i]: 1.0 / n for i in range(n)}}
        return self.step(state), info


def make(env_id, config=None, device='cuda:0'):
    """rlcard.make (envs/registration.py:77-89)."""
    return Env(env_id, config, device=device)
