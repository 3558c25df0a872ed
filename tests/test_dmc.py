"""The DMC actor data path (rlc_dmc_collect / rlcard_b200.dmc.DMCCollector) against the reference's own ``act``
loop (tests/golden/dmc_<game>.npz, recorded by tests/golden/make_dmc_golden.py from the live reference)."""
import os

import numpy as np
import pytest

import oracle
from dmc_util import feature_fn, fold_trajectory

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GAMES = ['doudizhu', 'leduc-holdem', 'uno', 'limit-holdem', 'scout']


def load(game):
    z = np.load(os.path.join(GOLDEN, 'dmc_%s.npz' % game.replace('-', '_')))
    return {k: z[k] for k in z.files}


def replay(env_reset, env_step, env_view, actions):
    """Drive an env through the recorded action stream episode by episode (env.run, env.py:120-169)."""
    obs, act, pl, done, pay = [], [], [], [], []
    k = 0
    while k < len(actions):
        env_reset()
        over = False
        while not over:
            o, p = env_view()
            obs.append(o); pl.append(p); act.append(int(actions[k]))
            over, py = env_step(int(actions[k]))
            done.append(over); pay.append(py)
            k += 1
    D = max(len(o) for o in obs)                   # seats may see different widths (doudizhu 790 / 901): zero pad
    obs = [np.concatenate([o, np.zeros(D - len(o), o.dtype)]) for o in obs]
    return (np.stack(obs)[:, None], np.asarray(act, np.int32)[:, None], np.asarray(pl, np.int32)[:, None],
            np.asarray(done, np.uint8)[:, None], np.stack(pay)[:, None])


def check_against_fixture(fx, rows, obs_dims):
    T = int(fx['T'])
    for p in range(int(fx['num_players'])):
        k = int(fx['filled_%d' % p])
        assert k > 0 and len(rows[p]) >= k * T
        got = rows[p][:k * T]
        want_state = fx['state_%d' % p].reshape(k * T, -1)
        d = obs_dims[p]
        assert want_state.shape[1] == d
        np.testing.assert_array_equal(np.stack([r[0][:d] for r in got]).astype(np.int8), want_state, err_msg='state p%d' % p)
        np.testing.assert_array_equal(np.stack([r[1] for r in got]), fx['action_%d' % p].reshape(k * T, -1), err_msg='action p%d' % p)
        np.testing.assert_array_equal(np.asarray([r[2] for r in got], np.float32), fx['target_%d' % p].reshape(-1))
        np.testing.assert_array_equal(np.asarray([r[3] for r in got], bool), fx['done_%d' % p].reshape(-1))
        np.testing.assert_array_equal(np.asarray([r[4] for r in got], np.float32), fx['episode_return_%d' % p].reshape(-1))


@pytest.mark.parametrize('game', GAMES)
def test_numpy_restatement_of_act_matches_reference(game):
    """oracle env replaying the recorded tape + actions -> fold_trajectory == the buffers the reference actor filled."""
    fx = load(game)
    env = oracle.OracleEnv(game)
    env.set_tape(fx['tape'])

    def step(a):
        env.step(a)
        return env.is_over(), np.asarray(env.payoffs(), np.float32) if env.is_over() else np.zeros(env.num_players, np.float32)
    tr = replay(env.reset, step, lambda: (env.obs(-1).copy(), env.player()), fx['actions'])
    # split into two windows to exercise the carry-over of an unfinished episode
    cut = len(fx['actions']) // 2
    f = feature_fn(game, env.num_actions)
    rows1, carry = fold_trajectory(*[x[:cut] for x in tr], f)
    rows2, carry = fold_trajectory(*[x[cut:] for x in tr], f, carry)
    assert all(len(c) == 0 for c in carry)
    rows = [a + b for a, b in zip(rows1, rows2)]
    check_against_fixture(fx, rows, env.obs_dim)
    assert env.tape_err() == 0


@pytest.mark.gpu
@pytest.mark.parametrize('window', [5, 1000])
@pytest.mark.parametrize('game', GAMES)
def test_cuda_collector_matches_reference_act(game, window):
    """CUDA env in replay mode + rlc_dmc_collect (in windows that cut through episodes) == the reference actor's buffers."""
    import torch
    import rlcard_b200
    from rlcard_b200.dmc import DMCCollector
    fx = load(game)
    env = rlcard_b200.VecEnv(game, 1, mode='replay', auto_reset=False)
    env.set_tape(fx['tape'][None, :])

    def step(a):
        env.step(torch.tensor([a], dtype=torch.int32, device='cuda'))
        return bool(env.done[0].item()), env.payoffs[0].cpu().numpy().copy()
    tr = replay(lambda: env.reset(), step, lambda: (env.obs[0].cpu().numpy().copy(), int(env.cur_player[0].item())), fx['actions'])
    col = DMCCollector(env, pool_rows=4096)
    L = tr[1].shape[0]
    for t0 in range(0, L, window):
        sl = slice(t0, min(L, t0 + window))
        traj = dict(obs=torch.from_numpy(tr[0][sl]).cuda().contiguous(), action=torch.from_numpy(tr[1][sl]).cuda().contiguous(),
                    player=torch.from_numpy(tr[2][sl]).cuda().contiguous(), done=torch.from_numpy(tr[3][sl]).cuda().contiguous(),
                    payoffs=torch.from_numpy(tr[4][sl].astype(np.float32)).cuda().contiguous())
        col.add(traj)
    sizes = col.sizes()
    assert int(col.open_len.sum().item()) == 0
    rows = []
    for p in range(env.num_players):
        n = sizes[p]
        st, ac = col.state[p][:n].cpu().numpy(), col.action[p][:n].cpu().numpy()
        tg, dn, er = col.target[p][:n].cpu().numpy(), col.done[p][:n].cpu().numpy(), col.episode_return[p][:n].cpu().numpy()
        rows.append([(st[i], ac[i], tg[i], bool(dn[i]), er[i]) for i in range(n)])
    check_against_fixture(fx, rows, env.obs_dims)
    # and the learner-facing batch (get_batch, utils.py:33-50): [T, B, ...] with the reference's dtypes
    T = int(fx['T'])
    b = col.get_batch(0, T, int(fx['filled_0']))
    assert b['state'].dtype == torch.int8 and b['action'].dtype == torch.int8 and b['done'].dtype == torch.bool
    np.testing.assert_array_equal(b['state'].cpu().numpy().transpose(1, 0, 2), fx['state_0'])
    np.testing.assert_array_equal(b['target'].cpu().numpy().T, fx['target_0'])
    assert col.sizes()[0] == sizes[0] - T * int(fx['filled_0'])


@pytest.mark.gpu
@pytest.mark.parametrize('game', ['doudizhu', 'leduc-holdem', 'scout'])
def test_cuda_collector_on_vectorised_rollouts(game):
    """Many envs, several rollout windows with auto reset: per position, the multiset of rows equals the numpy
    restatement applied to the same trajectories."""
    import torch
    import rlcard_b200
    from rlcard_b200.dmc import DMCCollector
    n, T, windows = 96, 40, 4
    env = rlcard_b200.VecEnv(game, n, seed=5)
    env.reset()
    col = DMCCollector(env, pool_rows=n * T * windows)
    f = feature_fn(game, env.num_actions)
    carry, want = None, [[] for _ in range(env.num_players)]
    for _ in range(windows):
        tr = col.collect_random(T)
        np_tr = [tr[k].cpu().numpy() for k in ('obs', 'action', 'player', 'done', 'payoffs')]
        rows, carry = fold_trajectory(*np_tr, f, carry)
        for p in range(env.num_players):
            want[p] += rows[p]
    sizes = col.sizes()
    for p in range(env.num_players):
        assert sizes[p] == len(want[p]) > 0
        key = lambda r: (r[0].tobytes(), np.asarray(r[1]).tobytes(), float(r[2]), bool(r[3]), float(r[4]))
        got = [(col.state[p][i].cpu().numpy(), col.action[p][i].cpu().numpy(), float(col.target[p][i]), bool(col.done[p][i]),
                float(col.episode_return[p][i])) for i in range(sizes[p])]
        assert sorted(map(key, got)) == sorted(map(key, want[p]))
    assert int(col.open_len.sum().item()) == sum(len(c) for c in carry)


@pytest.mark.gpu
@pytest.mark.parametrize('game', ['doudizhu', 'uno', 'leduc-holdem'])
def test_legal_ids_features_and_dmc_policy(game):
    """rlc_legal_ids / rlc_action_features against numpy on the env's masks, and DMCPolicy (greedy) against a numpy
    restatement of DMCAgent.step (model.py:60-110): argmax over the legal actions in ascending id order."""
    import torch
    import rlcard_b200
    from rlcard_b200.dmc import DMCCollector, DMCPolicy, action_features, legal_ids
    n = 200
    env = rlcard_b200.VecEnv(game, n, seed=17)
    env.reset()
    env.rollout_random(23)                                            # some mid-game states
    obs, mask, cur, _, _ = env.get_state(None)
    ids, count = legal_ids(env)
    m = mask.cpu().numpy()
    if env.mask_bitpacked:
        m = np.unpackbits(m.view(np.uint8), axis=1, bitorder='little')[:, :env.num_actions]
    f = feature_fn(game, env.num_actions)
    ids_np, cnt_np = ids.cpu().numpy(), count.cpu().numpy()
    feats = action_features(env, ids).cpu().numpy()
    for i in range(n):
        want = np.nonzero(m[i])[0]
        assert cnt_np[i] == len(want)
        assert ids_np[i, :len(want)].tolist() == want.tolist() and (ids_np[i, len(want):] == -1).all()
        for k, a in enumerate(want[:5]):
            np.testing.assert_array_equal(feats[i, k], f(int(a)))
        assert not feats[i, len(want):].any()
    F = feats.shape[-1]
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    ws = [(torch.randn(env.obs_dims[p], device='cuda', generator=g), torch.randn(F, device='cuda', generator=g)) for p in range(env.num_players)]
    nets = [lambda o, a, w=w: o @ w[0] + a @ w[1] for w in ws]
    acts = DMCPolicy(env, nets, exp_epsilon=0.0)(obs, mask, cur).cpu().numpy()
    o_np, cur_np = obs.cpu().numpy().astype(np.float32), cur.cpu().numpy()
    for i in range(n):
        p = cur_np[i]
        want = np.nonzero(m[i])[0]
        w0, w1 = ws[p][0].cpu().numpy(), ws[p][1].cpu().numpy()
        vals = np.array([o_np[i, :env.obs_dims[p]] @ w0 + f(int(a)).astype(np.float32) @ w1 for a in want])
        best = vals.max()
        assert acts[i] in want
        assert vals[list(want).index(acts[i])] >= best - 1e-3 * max(1.0, abs(best))     # argmax up to float reassociation
    # the policy drives the env and the collector end to end
    col = DMCCollector(env, pool_rows=n * 64)
    traj = {k: [] for k in ('obs', 'action', 'player', 'done', 'payoffs')}
    pol = DMCPolicy(env, nets, exp_epsilon=0.1)
    for _ in range(30):
        a = pol(env.obs, env.mask, env.cur_player)
        traj['obs'].append(env.obs.clone()); traj['player'].append(env.cur_player.clone()); traj['action'].append(a.int())
        env.step(a)
        traj['done'].append(env.done.clone()); traj['payoffs'].append(env.payoffs.clone())
    env.check_errors()                                                # every chosen action was legal
    col.add({k: torch.stack(v).contiguous() for k, v in traj.items()})
    assert sum(col.sizes()) + int(col.open_len.sum().item()) == n * 30
