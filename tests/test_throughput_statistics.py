"""Throughput (Philox) mode takes distribution-preserving shortcuts (partial Fisher-Yates deals, Blackjack without a
shuffle, UNO piles as multisets, chained small draws).  Bit-exactness with the reference is only defined in the replay
modes, so here the *distributions* are compared: uniform-random play under the Philox spec vs uniform-random play with
the reference's own chance stream (np.random.RandomState through the oracle), on episode length and payoffs."""
import numpy as np
import pytest

import oracle

CASES = {  # game: (envs, steps per env in throughput mode, episodes in reference-chance mode)
    'blackjack': (512, 200, 12000), 'leduc-holdem': (512, 200, 8000), 'limit-holdem': (512, 200, 6000),
    'no-limit-holdem': (512, 200, 6000), 'uno': (256, 600, 700), 'doudizhu': (64, 400, 160), 'scout': (64, 800, 120),
}


def philox_episodes(game, n, T, seed):
    tr = oracle.OracleVec(game, n, seed).rollout(T, want_obs=False, nthreads=8)
    lens, pay0 = [], []
    done, pay = tr['done'], tr['payoffs']
    for e in range(n):
        idx = np.nonzero(done[:, e])[0]
        prev = -1
        for t in idx:
            if prev >= 0:                                   # skip the first (possibly partial) episode of the window
                lens.append(t - prev); pay0.append(pay[t, e, 0])
            prev = t
    return np.asarray(lens, np.float64), np.asarray(pay0, np.float64)


def reference_chance_episodes(game, episodes, seed):
    env = oracle.OracleEnv(game)
    env.seed(seed)
    rng = np.random.RandomState(seed + 1)
    lens, pay0 = [], []
    for _ in range(episodes):
        env.reset()
        k = 0
        while not env.is_over():
            ids = np.nonzero(env.legal_mask())[0]
            env.step(int(ids[rng.randint(len(ids))]))
            k += 1
        lens.append(k); pay0.append(env.payoffs()[0])
    return np.asarray(lens, np.float64), np.asarray(pay0, np.float64)


def z(a, b):
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    return abs(a.mean() - b.mean()) / max(se, 1e-12)


@pytest.mark.parametrize('game', list(CASES))
def test_philox_mode_matches_reference_chance_in_distribution(game):
    n, T, episodes = CASES[game]
    la, pa = philox_episodes(game, n, T, seed=2024)
    lb, pb = reference_chance_episodes(game, episodes, seed=77)
    assert len(la) > 0.5 * episodes
    assert z(la, lb) < 5.0, ('episode length', la.mean(), lb.mean())
    assert z(pa, pb) < 5.0, ('payoff of seat 0', pa.mean(), pb.mean())
    assert z(la ** 2, lb ** 2) < 5.0, ('second moment of the episode length', (la ** 2).mean(), (lb ** 2).mean())
    assert z(pa ** 2, pb ** 2) < 5.0, ('second moment of the payoff', (pa ** 2).mean(), (pb ** 2).mean())
