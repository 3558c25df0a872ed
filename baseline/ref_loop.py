"""The reference's own CPU rollout loop, timed: ``env.run`` with ``RandomAgent``s on every seat
(the body of /root/reference/examples/run_random.py:23-27), one spawned process per host core
(the reference's multiprocessing model: spawn actors each owning an env seeded ``env.seed(i)``,
/root/reference/rlcard/agents/dmc_agent/trainer.py:223,269-276).

The UNMODIFIED reference package is installed by ``install()`` into the git-ignored ``baseline/_ref/``
(``pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>``; build container
only) and imported from there; nothing of this repo's engine is on that path.  ``termcolor`` (a hard
import of rlcard/games/uno/card.py:1, absent from the offline wheelhouse) is replaced by a 2-line shim.
Used only by bench.py (``cpu_baseline`` and ``--impl reference``).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
REFERENCE_SRC = '/root/reference'
GAMES = ('blackjack', 'leduc-holdem', 'limit-holdem', 'uno', 'doudizhu', 'scout', 'no-limit-holdem')


def installed():
    return os.path.isfile(os.path.join(REF_DIR, 'rlcard', '__init__.py'))


def install(force=False):
    """Build container only: install the reference into baseline/_ref (travels to the GPU box with the snapshot)."""
    if installed() and not force:
        return 'present'
    if not os.path.isdir(os.path.join(REFERENCE_SRC, 'rlcard')):
        return 'no reference source here'
    how = 'pip'
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'reference')            # /root/reference is read-only: build from a copy
        shutil.copytree(REFERENCE_SRC, src, ignore=shutil.ignore_patterns('.git', 'web', 'docs', '__pycache__'))
        shutil.rmtree(REF_DIR, ignore_errors=True)
        cmd = [sys.executable, '-m', 'pip', 'install', '-q', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', REF_DIR, src]
        rc = subprocess.call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
        if rc != 0 or not installed():                  # plain copy of the package (same files)
            how = 'copy'
            shutil.rmtree(REF_DIR, ignore_errors=True)
            os.makedirs(REF_DIR, exist_ok=True)
            shutil.copytree(os.path.join(REFERENCE_SRC, 'rlcard'), os.path.join(REF_DIR, 'rlcard'),
                            ignore=shutil.ignore_patterns('__pycache__'))
    with open(os.path.join(REF_DIR, 'termcolor.py'), 'w') as f:
        f.write('def colored(text, *args, **kwargs):\n    return text\n')
    # DouDizhu unzips its rule tables into its own package directory on first import: do it now
    subprocess.check_call([sys.executable, '-c', 'import sys; sys.path.insert(0, %r); import rlcard; rlcard.make("doudizhu")' % REF_DIR],
                          stdout=subprocess.DEVNULL)
    return how


def _worker(game, seed, seconds, is_training, q):
    sys.path.insert(0, REF_DIR)
    import numpy as np
    import rlcard
    from rlcard.agents import RandomAgent
    assert os.path.abspath(rlcard.__file__).startswith(REF_DIR), rlcard.__file__
    env = rlcard.make(game, config={'seed': seed})
    np.random.seed(seed)
    env.set_agents([RandomAgent(num_actions=env.num_actions) for _ in range(env.num_players)])
    env.run(is_training=is_training)                      # warm-up episode
    s0, episodes, t0 = env.timestep, 0, time.perf_counter()
    while True:
        env.run(is_training=is_training)
        episodes += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    q.put((env.timestep - s0, episodes, dt))


def run(game, seconds=10.0, procs=None, is_training=False):
    """-> dict(value = sum over processes of env-steps / s, ...).  One step = one Env.step call."""
    import multiprocessing as mp
    if not installed():
        raise RuntimeError('baseline/_ref is missing: run __graft_entry__.build() in the build container first')
    procs = procs or (os.cpu_count() or 1)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(game, i, seconds, is_training, q)) for i in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get(timeout=seconds * 20 + 300) for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    episodes = sum(r[1] for r in res)
    rate = sum(r[0] / r[2] for r in res)
    return {'value': rate, 'steps': steps, 'episodes': episodes, 'episodes_per_s': sum(r[1] / r[2] for r in res),
            'mean_episode_len': steps / max(1, episodes), 'procs': procs, 'seconds': seconds, 'wall_s': wall,
            'is_training': bool(is_training)}


if __name__ == '__main__':
    print(install())
    for g in (sys.argv[1:] or GAMES):
        print(g, run(g, seconds=3.0))
