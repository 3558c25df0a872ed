#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched simulator (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--game leduc-holdem] [--envs E] [--rollout-len T]
  python bench.py --impl reference ...      # the CPU arm: the reference's own Python loop on the host cores

One bench "step" = one pass of the hot path over the whole batch: ONE rlc_rollout_random launch that
advances each of the E envs of this GPU by T env-steps (uniform-random legal actions, auto reset) and
writes the full trajectory (obs + legal mask + action + player + done + payoffs) to HBM.  The
trajectory of one step (E*T*B bytes, hundreds of MB) is larger than the 126 MB L2, so every timed
iteration streams to DRAM.  N>1: one process per GPU (torchrun), envs sharded by global env id, no
collective on the step path; the only reduction is the end-of-run episode statistics.

The headline line is BASELINE.json config 2 (Leduc, 65 536 envs per GPU).  Unless --game is given, the
same JSON line carries a ``configs`` array with the same measurements (value, roofline, e2e, cpu_baseline)
for configs 3-5: Limit Hold'em, UNO, DouDizhu, Scout.

CPU arm: ``cpu_baseline`` (kind "reference") is the UNMODIFIED reference (baseline/_ref, installed by
__graft_entry__.build() in the build container) running ``env.run`` with RandomAgents, one spawned process per
host core, in this same invocation; ``cpu_baseline_port`` is the multi-threaded C oracle port (kind "port").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_ENVS = {'blackjack': 65536, 'leduc-holdem': 65536, 'limit-holdem': 16384, 'uno': 16384,
                'doudizhu': 8192, 'scout': 8192, 'no-limit-holdem': 16384}
HEADLINE = 'leduc-holdem'
EXTRA_CONFIGS = ('limit-holdem', 'uno', 'doudizhu', 'scout')       # BASELINE.json configs 3, 4, 5, 5
METRIC = 'env steps/sec (random policy, obs+mask)'
UNIT = 'env-steps/s'


def workload_name(game, envs, obs_dtype, obs_stride, num_actions):
    return '%s, %d envs per GPU, obs %s[%d] + legal mask u8[%d] + action/player/done/payoffs per env-step' % (
        game, envs, obs_dtype, obs_stride, num_actions)


def algorithmic_bytes_per_env_step(info, obs_elem, T):
    """B = O + M + 4 (action) + 4 (player) + 1 (done) + 4P (payoffs) per env-step, plus the
    packed state read+written once per launch (amortised over the T steps of the launch)."""
    O = info.obs_stride * obs_elem
    M = info.mask_words * 4 if info.mask_bitpacked else info.num_actions
    per_step = O + M + 4 + 4 + 1 + 4 * info.num_players
    state = 2 * 4 * info.state_words_philox      # the words the throughput kernels read + write per env and launch
    return per_step + state / float(T), per_step, state


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksEventReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksEventReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksEventReasonSwPowerCap: 'sw_power_cap'} if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else {
                 nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ------------------------------------------------------------------------------------------------ CPU arms
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for ln in open('/proc/cpuinfo'):
            if ln.startswith('model name'):
                return ln.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def cpu_port_run(game, envs, T, seed, threads, budget_s=12.0):
    """The oracle (C port of the reference algorithm) on the host cores: bounded sample of the same
    workload (same Philox chance/policy, trajectories written to host memory)."""
    import numpy as np
    import oracle
    n = min(envs, 8192)
    v = oracle.OracleVec(game, n, seed)
    L = v.L
    P, A = v.num_players, v.num_actions
    obs = np.zeros((T, n, v.obs_stride), np.float32)
    mask = np.zeros((T, n, A), np.uint8)
    act = np.zeros((T, n), np.int32); pl = np.zeros((T, n), np.int32)
    done = np.zeros((T, n), np.uint8); pay = np.zeros((T, n, P), np.float32)
    args = (v.h, T, obs.ctypes.data, v.obs_stride, mask.ctypes.data, act.ctypes.data, pl.ctypes.data,
            done.ctypes.data, pay.ctypes.data, threads)
    L.orc_envs_rollout(*args)                      # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        steps += L.orc_envs_rollout(*args)
        dt = time.perf_counter() - t0
        if dt > budget_s or steps >= 40_000_000_000:
            break
    return steps / dt, dt, steps, n


def cpu_port_baseline(game, envs, T, seed, budget_s):
    threads = host_cores()
    v, dt, steps, n = cpu_port_run(game, envs, T, seed, threads, budget_s)
    return {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': '%d env-steps in %.1f s: %d of the %d envs x %d-step launches of the multi-threaded C oracle port, '
                      'same Philox chance + policy, obs f32 + mask to host RAM' % (steps, dt, n, envs, T)}


def python_reference_baseline(game, seconds):
    """The reference itself (baseline/_ref): env.run + RandomAgent, one spawned process per host core."""
    from baseline import ref_loop
    if not ref_loop.installed():
        return None
    procs = host_cores()
    r = ref_loop.run(game, seconds=seconds, procs=procs, is_training=False)
    return {'value': r['value'], 'unit': UNIT, 'cores': procs, 'kind': 'reference', 'language': 'python',
            'cpu': cpu_model(),
            'episodes_per_s': r['episodes_per_s'], 'mean_episode_len': r['mean_episode_len'],
            'sample': '%d env-steps (%d episodes) in %.1f s per process: the unmodified reference from baseline/_ref, '
                      'rlcard.make(%r, seed=i) + env.run(is_training=False) with RandomAgents on every seat '
                      '(examples/run_random.py:23-27), %d spawned processes (one per host core)' % (
                          r['steps'], r['episodes'], seconds, game, procs)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its Python env.run loop with
    RandomAgents) on all host cores of this box; falls back to the C oracle port only when baseline/_ref did
    not travel."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import oracle
    oracle.build()
    T = args.rollout_len
    t_all = time.perf_counter()
    budget = min(30.0, max(6.0, 0.5 * args.steps))

    def one(game, envs, seconds):
        cb = python_reference_baseline(game, seconds)
        if cb is None:
            cb = cpu_port_baseline(game, envs, T, args.seed, seconds)
            cb['note'] = 'baseline/_ref (the Python reference) is not installed here: C oracle port instead'
        return cb

    cb = one(args.game, args.envs, budget)
    value = cb['value']
    np_, na, od = oracle.info(args.game)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * budget / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64' if args.game != 'scout' else 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.game, args.envs, 'float32' if args.game == 'scout' else 'uint8',
                                                 912 if args.game == 'doudizhu' else max(od), na),
                       'envs_per_gpu': args.envs, 'env_steps_per_launch_per_env': T,
                       'sample': 'per-env-step rate of the reference loop (one env per process, episodes back to back); a reference '
                                 '"step" is a %.2f s slice of that loop on every host core = a bounded sample (about %d env-steps) '
                                 'of the %d x %d env-steps of one bench step' % (budget / args.steps, int(value * budget / args.steps), args.envs, T)},
            'cpu_baseline': cb,
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if args.with_configs:
        line['configs'] = []
        for g in EXTRA_CONFIGS:
            c = one(g, DEFAULT_ENVS[g], min(budget, 6.0))
            line['configs'].append({'game': g, 'value': c['value'], 'unit': UNIT, 'cpu_baseline': c,
                                    'e2e': {'value': c['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
    line['wall_s'] = time.perf_counter() - t_all
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ NUMA binding
def bind_to_gpu_numa_node(local_rank, world):
    """Pin this rank (and therefore the pinned buffers it allocates, first touch) to the CPUs of the GPU's NUMA
    node; ranks sharing a node split its CPUs.  Returns a description for the JSON line."""
    info = {'bound': False}
    try:
        import pynvml
        pynvml.nvmlInit()
        cvd = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [x for x in cvd.split(',') if x.strip().isdigit()]
        idx = int(ids[local_rank]) if len(ids) > local_rank else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        try:
            mask = pynvml.nvmlDeviceGetCpuAffinityWithinScope(h, words, pynvml.NVML_AFFINITY_SCOPE_NODE)
        except Exception:
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = sorted(c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed] or allowed
        info['gpu_node_cpus'] = '%d-%d (%d)' % (cpus[0], cpus[-1], len(cpus))
        if world > 1 and len(cpus) >= world:
            # ranks whose GPUs share this CPU set take disjoint slices of it
            per = len(cpus) // world
            mine = cpus[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            info.update(bound=True, cpus='%d-%d' % (mine[0], mine[-1]))
        else:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus='%d-%d' % (cpus[0], cpus[-1]))
    except Exception as e:          # binding is an optimisation, never a failure
        info['error'] = repr(e)[:120]
    return info


# ------------------------------------------------------------------------------------------------ GPU arm
def bench_game(game, args, ctx, envs=None, steps=None, e2e_steps=None, headline=True):
    """All measurements of one game on this rank's GPU; rank 0 gets the assembled dict (others None)."""
    import numpy as np
    import torch
    import rlcard_b200
    from rlcard_b200 import _lib
    dist, world, rank, dev = ctx['dist'], ctx['world'], ctx['rank'], ctx['dev']
    info = rlcard_b200.game_info(game)
    obs_dtype = args.obs_dtype or ('float32' if info.obs_native_dtype == _lib.DTYPE_F32 else 'uint8')
    odt = torch.float32 if obs_dtype == 'float32' else torch.uint8
    E = envs or DEFAULT_ENVS[game]
    T, K, W = args.rollout_len, steps or args.steps, args.warmup
    L = rlcard_b200.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    env = rlcard_b200.VecEnv(game, E, device=dev, seed=args.seed, env_id_base=rank * E, obs_dtype=odt)
    env.reset()
    traj = env.alloc_trajectory(T)
    for _ in range(W):
        env.rollout_random(T, out=traj)
    barrier()
    sampler = ClockSampler(ctx['nvml_index'])
    sampler.start()
    launches0 = L.rlc_launch_count()
    stream = torch.cuda.current_stream(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(K):
        env.rollout_random(T, out=traj)
        ev[k + 1].record(stream)
    barrier()
    launches = L.rlc_launch_count() - launches0
    elapsed_ms = ev[0].elapsed_time(ev[K])
    kernel_ms = sum(ev[k].elapsed_time(ev[k + 1]) for k in range(K)) / K
    # keep the GPU under the same load a little longer so the clock record has enough samples
    t_end = time.perf_counter() + (0.6 if headline else 0.25)
    while time.perf_counter() < t_end:
        env.rollout_random(T, out=traj)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    env.check_errors()

    # episode statistics: the only cross-GPU reduction of the whole path
    done = traj['done'].bool()
    stats = torch.cat([(traj['payoffs'] * done.unsqueeze(-1)).sum((0, 1)).double(),
                       done.sum().double().reshape(1), torch.tensor([float(E * T)], device=dev, dtype=torch.float64)])
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    value = world * E * T * K / (elapsed_ms * 1e-3)
    del traj, env

    # ---- end to end through the public API with HOST buffers (VecEnv.rollout_random_host): every bench step
    # uploads the packed env state from pinned host memory, runs the same T-step random rollout in chunks and
    # streams the whole trajectory (obs + mask + action + player + done + payoffs) into pinned host memory
    # (chunk c copies D2H on a second stream while chunk c+1 is simulated), then reads the state back.
    e2e = None
    Ke = args.e2e_steps if e2e_steps is None else e2e_steps
    if Ke > 0:
        env2 = rlcard_b200.VecEnv(game, E, device=dev, seed=args.seed + 1, env_id_base=rank * E, obs_dtype=odt)
        env2.reset()
        h_traj = env2.alloc_host_trajectory(T)
        h_state = torch.empty(env2.state.shape, dtype=env2.state.dtype).pin_memory()
        h_state.copy_(env2.state)
        torch.cuda.synchronize(dev)
        for _ in range(min(W, 3) if headline else 1):
            env2.rollout_random_host(T, h_traj, chunk=args.e2e_chunk, host_state=h_state)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            env2.rollout_random_host(T, h_traj, chunk=args.e2e_chunk, host_state=h_state)
        torch.cuda.synchronize(dev)
        mine = time.perf_counter() - t0
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        per_rank = torch.zeros(world, device=dev, dtype=torch.float64)
        state_bytes = h_state.numel() * h_state.element_size()
        d2h = sum(x.numel() * x.element_size() for x in h_traj.values()) + state_bytes
        per_rank[rank] = (d2h + state_bytes) * Ke / mine / 1e9
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        e2e = {'value': world * E * T * Ke / float(dt.item()), 'unit': UNIT, 'h2d_bytes_per_step': state_bytes,
               'd2h_bytes_per_step': d2h, 'steps': Ke, 'ms_per_step': 1e3 * float(dt.item()) / Ke,
               'pcie_gbs': (d2h + state_bytes) * Ke / float(dt.item()) / 1e9,
               'pcie_gbs_per_rank': [round(float(x), 2) for x in per_rank.tolist()],
               'what': 'VecEnv.rollout_random_host: H2D packed state (pinned) -> %d-step chunks of rlc_rollout_random -> D2H of '
                       'the full trajectory into pinned host memory overlapped on a copy stream -> D2H state; host wall '
                       'clock, max over ranks' % args.e2e_chunk}
        env2.check_errors()
        assert h_traj['done'].numpy().any(), 'e2e trajectory did not reach the host'
        # ---- the same leg with the compact wire format (Leduc 4 B, Limit 12 B, UNO 28 B, DouDizhu 132 B, Scout 80 B per env-step): the chunk
        # is re-encoded on the device, only the records cross PCIe; the consumer expands rows on demand (compact.expand)
        if env2.compact_words():
            h_pack = env2.alloc_host_compact(T)
            for _ in range(min(W, 3) if headline else 1):
                env2.rollout_random_host(T, h_pack, chunk=args.e2e_chunk, host_state=h_state, compact=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(Ke):
                env2.rollout_random_host(T, h_pack, chunk=args.e2e_chunk, host_state=h_state, compact=True)
            torch.cuda.synchronize(dev)
            barrier()
            dtc = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
            d2hc = h_pack.numel() * 4 + state_bytes
            e2e['compact'] = {'value': world * E * T * Ke / float(dtc.item()), 'unit': UNIT, 'h2d_bytes_per_step': state_bytes,
                              'd2h_bytes_per_step': d2hc, 'ms_per_step': 1e3 * float(dtc.item()) / Ke,
                              'pcie_gbs': (d2hc + state_bytes) * Ke / float(dtc.item()) / 1e9,
                              'what': 'rollout_random_host(compact=True): %d-byte records per env-step re-encoded on the device '
                                      '(rlc_compact_trajectory), expanded lazily on the host; expansion is not in the timed region'
                                      % (4 * env2.compact_words())}
            assert (h_pack.numpy()[..., -1] >> 19 & 1).any(), 'compact e2e trajectory did not reach the host'
            del h_pack
        del h_traj, env2, h_state

    # ---- secondary: the per-step API driven by a host-resident agent (numpy uniform-legal policy), one
    # synchronous H2D(actions) -> rlc_step -> D2H(obs, mask, player, done, payoffs) round trip per env-step
    e2e_step = None
    if headline and info.num_actions == 4 and args.e2e_step_api_steps > 0:
        Ks = args.e2e_step_api_steps
        env2 = rlcard_b200.VecEnv(game, E, device=dev, seed=args.seed + 2, env_id_base=rank * E, obs_dtype=odt)
        h_act = torch.zeros(E, dtype=torch.int32).pin_memory()
        h_buf, hv = env2.alloc_host_step()
        h_mask = hv['mask']
        # k-th legal action LUT: code = 4 mask bits, r uniform in [0,12) (12 = lcm(1..4)) -> exact uniform
        lut = np.zeros((16, 12), np.int32)
        for code in range(1, 16):
            ids = [a for a in range(4) if (code >> a) & 1]
            for r in range(12):
                lut[code, r] = ids[r % len(ids)]
        rng = np.random.default_rng(args.seed)
        rnd = rng.integers(0, 12, size=(Ks + W, E), dtype=np.int64)
        m32 = h_mask.numpy().view(np.uint32).reshape(E)

        def host_step(i):
            code = ((m32.astype(np.uint64) * np.uint64(0x01020408)) >> np.uint64(24)) & np.uint64(15)
            np.take(lut.reshape(-1), code.astype(np.int64) * 12 + rnd[i], out=h_act.numpy())
            env2.step_host(h_act, h_buf)

        env2.reset()
        h_mask.copy_(env2.mask); torch.cuda.synchronize(dev)
        for i in range(W):
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ks):
            host_step(W + i)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_step = {'value': world * E * Ks / float(dt.item()), 'unit': UNIT, 'h2d_bytes_per_step': h_act.numel() * 4,
                    'd2h_bytes_per_step': h_buf.numel(), 'steps': Ks,
                    'what': 'VecEnv.step_host per env-step: numpy uniform-legal policy on host -> H2D actions (pinned) -> '
                            'rlc_step -> one D2H of obs+mask+player+done+payoffs (pinned), synchronous round trip'}
        env2.check_errors()
        del env2

    # ---- optional: the DMC actor data path (config 5): rollout window + rlc_dmc_collect into per-position pools
    dmc = None
    if args.dmc_steps > 0 and (headline or info.threads_per_env == 32):
        from rlcard_b200.dmc import DMCCollector
        env3 = rlcard_b200.VecEnv(game, E, device=dev, seed=args.seed + 3, env_id_base=rank * E, obs_dtype=odt)
        env3.reset()
        col = DMCCollector(env3, pool_rows=E * T + E * 8)
        traj3 = env3.alloc_trajectory(T)
        for _ in range(W):
            col.collect_random(T, traj3); col.count.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rows = 0
        for _ in range(args.dmc_steps):
            col.collect_random(T, traj3)
            rows += sum(col.sizes()); col.count.zero_()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        dmc = {'value': world * E * T * args.dmc_steps / (ms * 1e-3), 'unit': UNIT, 'steps': args.dmc_steps,
               'rows_per_step': rows / args.dmc_steps, 'ms_per_step': ms / args.dmc_steps,
               'what': 'rlc_rollout_random (T=%d) + rlc_dmc_collect into per-position (state int8, action feature, target, '
                       'done, episode_return) pools, device resident' % T}
        del env3, col, traj3
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'MEASURED_PEAKS.json hbm_gbs (measured)'
    else:
        peak, peak_src = 6650.0, 'B200_PROFILING.md fallback'
    b_step, per_step, state_b = algorithmic_bytes_per_env_step(info, 4 if odt == torch.float32 else 1, T)
    achieved = E * T * b_step / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('%s:%s:%d:%d' % (game, obs_dtype, E, T))
    kernel = ('k_rollout_leduc_fsm' if game == 'leduc-holdem' else 'k_rollout_limit_pipe' if game == 'limit-holdem' else
              ('k_wrollout<%s>' if info.threads_per_env == 32 else 'k_rollout<%s>') % game)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': elapsed_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if odt == torch.float32 else 'u8', 'data': 'synthetic',
        'config': {'workload': workload_name(game, E, obs_dtype, info.obs_stride, info.num_actions),
                   'envs_per_gpu': E, 'env_steps_per_launch_per_env': T, 'policy': 'uniform-random legal (Philox, on device)',
                   'chance': 'Philox4x32-10 keyed (seed, global env id, env-step index)', 'auto_reset': True,
                   'l2': 'trajectory written per launch = %.0f MB > 126 MB L2' % (E * T * per_step / 1e6)},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'kernel': kernel, 'kernel_ms': kernel_ms,
                     'bytes_per_env_step': b_step, 'peak_source': peak_src},
        'episodes_last_launch': float(stats[info.num_players].item()),
        'mean_payoff_per_seat': [float(x) / max(1.0, float(stats[info.num_players].item())) for x in stats[:info.num_players]],
    }
    if e2e:
        line['e2e'] = e2e
    if e2e_step:
        line['e2e_step_api'] = e2e_step
    if dmc:
        line['dmc'] = dmc
    if not args.no_cpu_baseline and world == 1:
        cb = python_reference_baseline(game, args.cpu_seconds if headline else min(args.cpu_seconds, 5.0))
        port = cpu_port_baseline(game, E, T, args.seed, 12.0 if headline else 3.0)
        if cb is not None:
            line['cpu_baseline'] = cb
            line['cpu_baseline_port'] = port
        else:
            port['note'] = 'baseline/_ref (the Python reference) is not installed here: C oracle port only'
            line['cpu_baseline'] = port
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--game', default=None, help='one game only (default: Leduc headline + configs array)')
    ap.add_argument('--envs', type=int, default=None, help='envs per GPU')
    ap.add_argument('--rollout-len', type=int, default=128, help='env-steps per env per launch (T)')
    ap.add_argument('--obs-dtype', default=None, choices=['uint8', 'float32'])
    ap.add_argument('--seed', type=int, default=20261018)
    ap.add_argument('--e2e-steps', type=int, default=10, help='bench steps (T-step host rollouts) of the e2e leg')
    ap.add_argument('--e2e-chunk', type=int, default=16, help='env-steps per rlc_rollout_random launch in the e2e leg')
    ap.add_argument('--e2e-step-api-steps', type=int, default=100, help='env-steps of the per-step host-agent leg')
    ap.add_argument('--dmc-steps', type=int, default=0, help='bench steps of the optional DMC-collector leg')
    ap.add_argument('--cpu-seconds', type=float, default=10.0, help='wall time per process of the Python reference loop')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='headline game only')
    ap.add_argument('--no-numa-bind', action='store_true')
    args = ap.parse_args()
    args.with_configs = args.game is None and not args.no_configs
    if args.game is None:
        args.game = HEADLINE
    if args.envs is None:
        args.envs = DEFAULT_ENVS[args.game]
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    numa = {'bound': False} if args.no_numa_bind else bind_to_gpu_numa_node(local, int(os.environ.get('LOCAL_WORLD_SIZE', world)))

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    cvd = os.environ.get('CUDA_VISIBLE_DEVICES', '')
    ids = [x for x in cvd.split(',') if x.strip().isdigit()]
    ctx = {'dist': dist, 'world': world, 'rank': rank, 'dev': dev,
           'nvml_index': int(ids[local]) if len(ids) > local else local}

    line = bench_game(args.game, args, ctx, envs=args.envs, headline=True)
    configs = []
    if args.with_configs:
        for g in EXTRA_CONFIGS:
            heavy = g in ('doudizhu', 'scout')
            c = bench_game(g, args, ctx, steps=min(args.steps, 20), e2e_steps=min(args.e2e_steps, 2 if heavy else 5), headline=False)
            if c is not None:
                c['game'] = g
                configs.append(c)
    if rank == 0:
        line['numa'] = numa
        if configs:
            line['configs'] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
