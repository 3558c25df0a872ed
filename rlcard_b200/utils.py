"""Callers of the path that live in rlcard/utils/utils.py: ``reorganize`` (:153-179) and ``tournament`` (:200-225),
for the single-env facade (same inputs / outputs as the reference) and, for ``tournament``, the vector form."""
import numpy as np


def reorganize(trajectories, payoffs):
    """Per-seat ``[s0, a0, s1, a1, ..., s_terminal]`` lists (Env.run) -> per-seat lists of
    ``[state, action, reward, next_state, done]``; the reward is the seat's payoff on its last transition, else 0."""
    out = []
    for seat, traj in enumerate(trajectories):
        n_transitions = (len(traj) - 1) // 2
        rows = []
        for k in range(n_transitions):
            final = k == n_transitions - 1
            state, action, next_state = traj[2 * k], traj[2 * k + 1], traj[2 * k + 2]
            rows.append([state, action, payoffs[seat] if final else 0, next_state, final])
        out.append(rows)
    return out


def tournament(env, num):
    """Mean payoff per seat over ``num`` games of ``env.run(is_training=False)`` (an env whose run() returns a list of
    payoff vectors contributes one game per entry)."""
    totals = [0 for _ in range(env.num_players)]
    games = 0
    while games < num:
        _, payoffs = env.run(is_training=False)
        batch = payoffs if isinstance(payoffs, list) else [payoffs]
        for p in batch:
            for seat in range(env.num_players):
                totals[seat] += p[seat]
            games += 1
    return [t / games for t in totals]


def vec_tournament(env, num_steps):
    """``tournament`` over a VecEnv with on-device random agents: mean payoff per seat over the episodes finished in
    ``num_steps`` rollout steps -> (list of means, number of games)."""
    tr = env.rollout_random(num_steps, out=env.alloc_trajectory(num_steps, obs=False, mask=False))
    done = tr['done'].bool()
    n = int(done.sum().item())
    tot = (tr['payoffs'] * done.unsqueeze(-1)).sum((0, 1)).double()
    return (tot / max(n, 1)).tolist(), n


def remove_illegal(action_probs, legal_actions):
    """rlcard/utils/utils.py:181-198: zero the illegal entries and renormalise (uniform over legal if all mass was illegal)."""
    probs = np.zeros(action_probs.shape[0])
    probs[legal_actions] = action_probs[legal_actions]
    s = probs.sum()
    if s == 0:
        probs[legal_actions] = 1 / len(legal_actions)
    else:
        probs /= s
    return probs
