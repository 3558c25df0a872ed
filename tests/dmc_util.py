"""numpy restatement of the reference DMC actor's bookkeeping (rlcard/agents/dmc_agent/utils.py:121-137) over a
trajectory window -- the checker for rlc_dmc_collect."""
import numpy as np


def fold_trajectory(obs, action, player, done, payoffs, feature, open_rows=None):
    """obs [L, n, D], action/player/done [L, n], payoffs [L, n, P]; feature(a) -> int8 vector.
    Returns (rows, open_rows): rows[p] = list of (state, action feature, target, done, episode_return) in the order
    the reference appends them, per env (rows of different envs are concatenated env by env); open_rows carries
    the decisions of episodes that have not ended (per env) into the next call."""
    L, n = action.shape
    P = payoffs.shape[-1]
    rows = [[] for _ in range(P)]
    open_rows = open_rows if open_rows is not None else [[] for _ in range(n)]
    for e in range(n):
        cur = open_rows[e]
        for t in range(L):
            cur.append((obs[t, e].copy(), int(action[t, e]), int(player[t, e])))
            if done[t, e]:
                pay = payoffs[t, e]
                for p in range(P):
                    mine = [r for r in cur if r[2] == p]
                    for k, (o, a, _) in enumerate(mine):
                        last = k == len(mine) - 1
                        rows[p].append((o, feature(a), float(pay[p]), last, float(pay[p]) if last else 0.0))
                cur = []
        open_rows[e] = cur
    return rows, open_rows


def feature_fn(game, num_actions):
    if game == 'doudizhu':
        from rlcard_b200 import doudizhu_table
        feats = doudizhu_table.load()['features']
        return lambda a: feats[a]
    eye = np.eye(num_actions, dtype=np.int8)
    return lambda a: eye[a]
