"""N>1 host path on CPU: 2 gloo ranks shard the global env-id range; their union equals one unsharded
run (Philox keyed by global env id) and the end-of-run statistics reduce correctly.  The per-rank
'device' results are produced by the CPU oracle here (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world_size, port, E, T, seed, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    import oracle
    from rlcard_b200 import distributed as D
    assert D.world() == (rank, world_size, rank)
    base = D.shard_env_id_base(E, rank)
    tr = oracle.OracleVec('leduc-holdem', E, seed, env0=base).rollout(T)
    stats = D.episode_stats(torch.from_numpy(tr['payoffs']), torch.from_numpy(tr['done']), E * T)
    local = stats.clone()
    D.reduce_stats(stats)
    slowest = D.reduce_max(10.0 + rank, 'cpu')
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), local=local.numpy(), total=stats.numpy(), slowest=slowest,
             action=tr['action'], obs=tr['obs'], done=tr['done'])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_unsharded_run(tmp_path):
    import oracle
    E, T, seed, W = 96, 24, 991, 2
    port = _free_port()
    mp.spawn(_worker, args=(W, port, E, T, seed, str(tmp_path)), nprocs=W, join=True)
    full = oracle.OracleVec('leduc-holdem', W * E, seed, env0=0).rollout(T)
    parts = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(W)]
    for k in ('action', 'obs', 'done'):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts], axis=1), full[k])
    total = parts[0]['total']
    np.testing.assert_array_equal(total, parts[1]['total'])
    np.testing.assert_allclose(total, parts[0]['local'] + parts[1]['local'])
    assert total[-1] == W * E * T and total[-2] == full['done'].sum()
    assert parts[0]['slowest'] == parts[1]['slowest'] == 11.0
