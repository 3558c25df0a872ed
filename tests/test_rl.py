"""run_rl.py's data path (TransitionCollector / rlc_rl_feed, and the host-side reorganize) against transitions produced
by the live reference's env.run(is_training=True) + reorganize (tests/golden/rl_<game>.npz, make_rl_golden.py)."""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GAMES = ['leduc-holdem', 'limit-holdem', 'uno', 'doudizhu', 'blackjack', 'scout', 'no-limit-holdem']


def load(game):
    z = np.load(os.path.join(GOLDEN, 'rl_%s.npz' % game.replace('-', '_')))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize('game', GAMES)
def test_host_reorganize_over_oracle_env_run(game):
    """Env.run (env.py:120-169) restated over the oracle env + rlcard_b200.utils.reorganize == the reference's rows."""
    from rlcard_b200.utils import reorganize
    fx = load(game)
    env = oracle.OracleEnv(game)
    env.set_tape(fx['tape'])
    P, A = env.num_players, env.num_actions
    rows = [[] for _ in range(P)]
    k, actions = 0, fx['actions']

    def state(seat):
        return {'obs': env.obs(seat).copy(), 'legal': env.legal_mask().copy()}
    while k < len(actions):
        traj = [[] for _ in range(P)]
        env.reset()
        pid = env.player()
        traj[pid].append(state(-1))
        while not env.is_over():
            a = int(actions[k]); k += 1
            env.step(a)
            traj[pid].append(a)
            pid = env.player()
            if not env.is_over():
                traj[pid].append(state(-1))
        for p in range(P):
            traj[p].append(state(p))
        for p, ts in enumerate(reorganize(traj, env.payoffs())):
            rows[p] += ts
    for p in range(P):
        d = env.obs_dim[p]
        assert len(rows[p]) == len(fx['action_%d' % p])
        legal = np.unpackbits(fx['next_legal_%d' % p], axis=1, bitorder='little')[:, :A]
        for i, (s, a, r, s2, done) in enumerate(rows[p]):
            tag = '%s seat %d row %d' % (game, p, i)
            np.testing.assert_array_equal(s['obs'].astype(np.float64), fx['state_%d' % p][i, :d].astype(np.float64), err_msg=tag)
            np.testing.assert_array_equal(s2['obs'].astype(np.float64), fx['next_state_%d' % p][i, :d].astype(np.float64), err_msg=tag)
            np.testing.assert_array_equal(s2['legal'], legal[i], err_msg=tag)
            assert a == fx['action_%d' % p][i] and float(r) == fx['reward_%d' % p][i] and bool(done) == bool(fx['done_%d' % p][i]), tag


@pytest.mark.gpu
@pytest.mark.parametrize('game', GAMES)
def test_cuda_transition_collector_matches_reference(game):
    import torch
    import rlcard_b200
    from rlcard_b200.rl import TransitionCollector
    fx = load(game)
    env = rlcard_b200.VecEnv(game, 1, mode='replay', auto_reset=False, terminal_obs=True)
    tape = np.concatenate([fx['tape'], np.zeros(512, np.uint8)])          # the collector deals once more after the last episode
    env.set_tape(tape[None, :])
    col = TransitionCollector(env, pool_rows=4096)
    col.reset()
    for a in fx['actions']:
        col.step(torch.tensor([int(a)], dtype=torch.int32, device='cuda'))
    assert not bool(col.pend_valid.any())
    A = env.num_actions
    for p in range(env.num_players):
        got = col.pop(p)
        d = env.obs_dims[p]
        n = len(fx['action_%d' % p])
        assert got['action'].shape[0] == n, (game, p)
        np.testing.assert_array_equal(got['state'].cpu().numpy().astype(np.float64), fx['state_%d' % p][:, :d].astype(np.float64))
        np.testing.assert_array_equal(got['next_state'].cpu().numpy().astype(np.float64), fx['next_state_%d' % p][:, :d].astype(np.float64))
        np.testing.assert_array_equal(got['action'].cpu().numpy(), fx['action_%d' % p])
        np.testing.assert_array_equal(got['reward'].cpu().numpy().astype(np.float64), fx['reward_%d' % p])
        np.testing.assert_array_equal(got['done'].cpu().numpy().astype(np.uint8), fx['done_%d' % p])
        m = got['next_mask'].cpu().numpy()
        if env.mask_bitpacked:
            m = np.unpackbits(m.view(np.uint8), axis=1, bitorder='little')[:, :A]
        np.testing.assert_array_equal(m, np.unpackbits(fx['next_legal_%d' % p], axis=1, bitorder='little')[:, :A])
    assert col.sizes() == [0] * env.num_players


@pytest.mark.gpu
def test_collector_run_with_batched_policy():
    """Many envs driven by a torch policy: per seat, transitions chain (next_state of one row is the state of the seat's
    next row inside an episode), rewards only on done rows, in half big blinds."""
    import torch
    import rlcard_b200
    from rlcard_b200.rl import TransitionCollector
    env = rlcard_b200.VecEnv('leduc-holdem', 512, seed=3, auto_reset=False, terminal_obs=True)
    col = TransitionCollector(env, pool_rows=512 * 64)
    col.reset()
    col.run(rlcard_b200.random_policy(), 40)
    t0, t1 = col.pop(0), col.pop(1)
    assert t0['action'].shape[0] > 2000 and t1['action'].shape[0] > 2000
    for t in (t0, t1):
        assert bool((t['reward'][~t['done']] == 0).all())
        assert bool(((t['reward'] * 2) == (t['reward'] * 2).round()).all())
        assert bool((t['next_mask'].sum(1) >= 1).all())


@pytest.mark.gpu
@pytest.mark.parametrize('window', [7, 4096])
@pytest.mark.parametrize('game', GAMES)
def test_fused_rollout_reorganize_matches_reference(game, window):
    """The reference's run_rl.py fixtures replayed through the FUSED path: rlc_rollout_random in replay mode (recorded
    chance tape + the recorded action ids as forced_actions, auto reset, terminal-state pool) followed by rlc_reorganize,
    in windows that cut through episodes.  Every transition row equals the reference's reorganize output."""
    import torch
    import rlcard_b200
    from rlcard_b200.rl import TransitionCollector
    fx = load(game)
    env = rlcard_b200.VecEnv(game, 1, mode='replay')
    tape = np.concatenate([fx['tape'], np.zeros(512, np.uint8)])          # the auto reset after the last episode deals once more
    env.set_tape(tape[None, :])
    env.reset()
    col = TransitionCollector(env, pool_rows=4096, fused=True)
    acts = torch.tensor(fx['actions'], dtype=torch.int32, device='cuda').view(-1, 1)
    for t0 in range(0, len(acts), window):
        chunk = acts[t0:t0 + window]
        pad = torch.full((window - len(chunk), 1), -1, dtype=torch.int32, device='cuda')    # idle cells
        col.collect_fused(window, actions=torch.cat([chunk, pad]).contiguous())
    assert not (int(env.err.item()) & ~4)                                  # 4 = the fixtures' deliberate illegal ids
    assert not bool(col.pend_valid.any())
    A = env.num_actions
    for p in range(env.num_players):
        got = col.pop(p)
        d = env.obs_dims[p]
        n = len(fx['action_%d' % p])
        assert got['action'].shape[0] == n, (game, p)
        np.testing.assert_array_equal(got['state'].cpu().numpy().astype(np.float64), fx['state_%d' % p][:, :d].astype(np.float64))
        np.testing.assert_array_equal(got['next_state'].cpu().numpy().astype(np.float64), fx['next_state_%d' % p][:, :d].astype(np.float64))
        np.testing.assert_array_equal(got['action'].cpu().numpy(), fx['action_%d' % p])
        np.testing.assert_array_equal(got['reward'].cpu().numpy().astype(np.float64), fx['reward_%d' % p])
        np.testing.assert_array_equal(got['done'].cpu().numpy().astype(np.uint8), fx['done_%d' % p])
        m = got['next_mask'].cpu().numpy()
        if env.mask_bitpacked:
            m = np.unpackbits(m.view(np.uint8), axis=1, bitorder='little')[:, :A]
        np.testing.assert_array_equal(m, np.unpackbits(fx['next_legal_%d' % p], axis=1, bitorder='little')[:, :A])


@pytest.mark.gpu
@pytest.mark.parametrize('game', GAMES)
def test_fused_reorganize_equals_numpy_restatement(game):
    """Throughput mode, many envs, three chained windows: the device rows of rlc_reorganize equal a numpy restatement of
    Env.run + reorganize (env.py:120-169, utils.py:153-179) applied to the same fused trajectory + terminal pool, per
    seat as multisets (pool order depends on warp scheduling)."""
    import torch
    import rlcard_b200
    from rlcard_b200.rl import TransitionCollector
    n, T, windows, seed = 64, 20, 3, 31
    env = rlcard_b200.VecEnv(game, n, seed=seed)
    env.reset()
    col = TransitionCollector(env, pool_rows=n * T * windows * 2, fused=True)
    P = env.num_players
    want = [[] for _ in range(P)]
    pend = [[None] * P for _ in range(n)]
    episodes = 0
    for _ in range(windows):
        w = {k: v.cpu().numpy() for k, v in col.collect_fused(T).items()}
        for i in range(n):
            for t in range(T):
                a, p = int(w['action'][t, i]), int(w['player'][t, i])
                o, m = w['obs'][t, i], w['mask'][t, i]
                if pend[i][p] is not None:
                    want[p].append((pend[i][p][0].tobytes(), pend[i][p][1], 0.0, o.tobytes(), m.tobytes(), False))
                pend[i][p] = (o.copy(), a)
                if w['done'][t, i]:
                    episodes += 1
                    r = int(w['terminal_row'][t, i])
                    assert r >= 0
                    for s in range(P):
                        if pend[i][s] is not None:
                            want[s].append((pend[i][s][0].tobytes(), pend[i][s][1], float(w['payoffs'][t, i, s]),
                                            w['terminal_obs'][r, s].tobytes(), w['terminal_mask'][r].tobytes(), True))
                        pend[i][s] = None
                else:
                    assert int(w['terminal_row'][t, i]) == -1
    env.check_errors()
    assert episodes > 0 or game in ('scout',)
    sizes = col.sizes()
    for p in range(P):
        assert sizes[p] == len(want[p])
        st, nx = col.state[p][:sizes[p]].cpu().numpy(), col.next_state[p][:sizes[p]].cpu().numpy()
        mk, ac = col.next_mask[p][:sizes[p]].cpu().numpy(), col.action[p][:sizes[p]].cpu().numpy()
        rw, dn = col.reward[p][:sizes[p]].cpu().numpy(), col.done[p][:sizes[p]].cpu().numpy()
        got = [(st[k].tobytes(), int(ac[k]), float(rw[k]), nx[k].tobytes(), mk[k].tobytes(), bool(dn[k])) for k in range(sizes[p])]
        assert sorted(got) == sorted(want[p]), (game, p)
    live = sum(1 for i in range(n) for s in range(P) if pend[i][s] is not None)
    assert int(col.pend_valid.sum().item()) == live
