"""GPU parity tests proper: the CUDA path, called through the C ABI (rlcard_b200.VecEnv / make),
against (1) fixtures recorded from the live reference and (2) the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

import oracle
import rlcard_b200
from replay_util import (ALL_FIXTURES, ALL_GAMES, DEEP_GAMES, check_record_deep, check_slot, deep_slot_records, fixture_game,
                         have_fixture, load_deep, load_fixture, slot_records, slot_tape)

pytestmark = pytest.mark.gpu


def built(game):
    return rlcard_b200.game_info(game).state_words > 0


GAMES = [g for g in ALL_GAMES if built(g)]
FIX_GAMES = [g for g in ALL_FIXTURES if have_fixture(g) and built(fixture_game(g))]


def to_np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize('chance', ['replay', 'mt19937'])
@pytest.mark.parametrize('obs_dtype', [torch.uint8, torch.float32])
@pytest.mark.parametrize('game', FIX_GAMES)
def test_replay_reference_tapes_vectorised(game, obs_dtype, chance):
    """Replay mode: all fixture slots side by side in one VecEnv, lock-step over the record streams;
    every obs / legal set / player / done / payoff must equal the reference's.  chance = 'replay' feeds the recorded
    draws, 'mt19937' only the slots' seeds (np.random.RandomState runs on the device, one generator per env)."""
    fx = load_fixture(game)
    game = fixture_game(game)
    if rlcard_b200.game_info(game).obs_native_dtype == 1 and obs_dtype == torch.uint8:
        pytest.skip('fractional obs')
    if chance == 'mt19937' and obs_dtype == torch.float32 and rlcard_b200.game_info(game).obs_native_dtype == 0:
        pytest.skip('covered by the uint8 case')
    S = len(fx['slot_seed'])
    recs = [slot_records(fx, s) for s in range(S)]
    L = max(len(slot_tape(fx, s)) for s in range(S))
    tape = np.zeros((S, L + 8), np.uint8)
    for s in range(S):
        t = slot_tape(fx, s); tape[s, :len(t)] = t
    env = rlcard_b200.VecEnv(game, S, mode=chance, obs_dtype=obs_dtype, auto_reset=False)
    if chance == 'replay':
        env.set_tape(tape)
    else:
        env.seed_mt19937([int(x) for x in fx['slot_seed']])
    checked = 0
    for tick in range(max(len(r) for r in recs)):
        live = [s for s in range(S) if tick < len(recs[s])]
        kinds = {s: int(fx['rec_kind'][recs[s][tick]]) for s in live}
        for kind in (0, 1, 2, 3):
            group = [s for s in live if kinds[s] == kind]
            if not group:
                continue
            if kind == 0:
                m = torch.zeros(S, dtype=torch.uint8); m[group] = 1
                env.reset(m.cuda())
            elif kind == 1:
                a = torch.full((S,), -1, dtype=torch.int32)
                for s in group:
                    a[s] = int(fx['rec_arg'][recs[s][tick]])
                env.step(a.cuda())
            elif kind == 2:
                seat = torch.zeros(S, dtype=torch.int32)
                for s in group:
                    seat[s] = int(fx['rec_arg'][recs[s][tick]])
                env.get_state(seat.cuda())
            else:
                env.get_state(None)
            obs, mask, cur, done, pay = (to_np(x) for x in (env.obs, env.mask, env.cur_player, env.done, env.payoffs))
            if env.mask_bitpacked:
                mask = np.unpackbits(mask.view(np.uint8), axis=1, bitorder='little')[:, :env.num_actions]
            for s in group:
                r = recs[s][tick]
                tag = '%s slot %d rec %d kind %d' % (game, s, r, kind)
                if kind == 3:
                    np.testing.assert_array_equal(pay[s].astype(np.float64), fx['rec_payoffs'][r], err_msg=tag)
                else:
                    d = int(fx['rec_obs_dim'][r])
                    np.testing.assert_array_equal(obs[s, :d].astype(np.float64), fx['rec_obs'][r, :d].astype(np.float64), err_msg=tag)
                    assert not obs[s, d:].any(), tag
                    np.testing.assert_array_equal(mask[s], fx['legal'][r], err_msg=tag)
                assert cur[s] == fx['rec_player'][r], tag
                assert bool(done[s]) == bool(fx['rec_done'][r]), tag
                checked += 1
    assert checked == len(fx['rec_slot'])
    err = to_np(env.err)
    assert not (err & 3).any()                      # tape never exhausted / out of range
    if chance == 'replay':
        pos = to_np(env.tape_pos)
        for s in range(S):
            assert pos[s] == len(slot_tape(fx, s))      # and consumed exactly


@pytest.mark.parametrize('chance', ['replay', 'mt19937'])
@pytest.mark.parametrize('game', [g for g in DEEP_GAMES if built(g)])
def test_replay_deep_reference_fixture(game, chance):
    """The deep fixtures (tests/golden/deep_<game>.npz: >= 11 000 reference episodes for the short games, 1 000+ UNO,
    500+ DouDizhu / Scout, stratified and crafted slots) replayed on the device, all slots side by side in lock step;
    obs rows and legal sets are compared through their digests, players / done / payoffs directly.  'replay' runs
    EVERY slot from its recorded draws; 'mt19937' every 4th seeded slot from the seed alone."""
    fx = load_deep(game)
    slots = list(range(len(fx['slot_seed'])))
    if chance == 'mt19937':
        slots = [s for s in slots[::4] if fx['slot_seed'][s] >= 0]
    S = len(slots)
    recs = [deep_slot_records(fx, s) for s in slots]
    tapes = [fx['tape'][fx['tape_off'][s]:fx['tape_off'][s + 1]] for s in slots]
    odt = torch.float32 if rlcard_b200.game_info(game).obs_native_dtype == 1 else torch.uint8
    env = rlcard_b200.VecEnv(game, S, mode=chance, obs_dtype=odt, auto_reset=False)
    if chance == 'replay':
        tape = np.zeros((S, max(len(t) for t in tapes) + 8), np.uint8)
        for i, t in enumerate(tapes):
            tape[i, :len(t)] = t
        env.set_tape(tape)
    else:
        env.seed_mt19937([int(fx['slot_seed'][s]) for s in slots])
    kind_all, arg_all = fx['rec_kind'], fx['rec_arg']
    checked = 0
    for tick in range(max(len(r) for r in recs)):
        live = [i for i in range(S) if tick < len(recs[i])]
        kinds = {i: int(kind_all[recs[i][tick]]) for i in live}
        for kind in (0, 1, 2, 3):
            group = [i for i in live if kinds[i] == kind]
            if not group:
                continue
            if kind == 0:
                m = torch.zeros(S, dtype=torch.uint8); m[group] = 1
                env.reset(m.cuda())
            elif kind == 1:
                a = torch.full((S,), -1, dtype=torch.int32)
                a[group] = torch.tensor([int(arg_all[recs[i][tick]]) for i in group], dtype=torch.int32)
                env.step(a.cuda())
            elif kind == 2:
                seat = torch.zeros(S, dtype=torch.int32)
                seat[group] = torch.tensor([int(arg_all[recs[i][tick]]) for i in group], dtype=torch.int32)
                env.get_state(seat.cuda())
            else:
                env.get_state(None)
            cur, done, pay = (to_np(x) for x in (env.cur_player, env.done, env.payoffs))
            if kind != 3:
                obs, mask = to_np(env.obs), to_np(env.mask)
                if env.mask_bitpacked:
                    mask = np.unpackbits(mask[group].view(np.uint8), axis=1, bitorder='little')[:, :env.num_actions]
                else:
                    mask = mask[group]
            for gi, i in enumerate(group):
                r = recs[i][tick]
                tag = '%s deep slot %d rec %d kind %d' % (game, slots[i], r, kind)
                if kind == 3:
                    np.testing.assert_array_equal(pay[i].astype(np.float64), fx['rec_payoffs'][r], err_msg=tag)
                else:
                    check_record_deep(fx, r, obs[i], mask[gi], tag)
                assert cur[i] == fx['rec_player'][r], tag
                assert bool(done[i]) == bool(fx['rec_done'][r]), tag
                checked += 1
    assert checked == sum(len(r) for r in recs)
    err = to_np(env.err)
    assert not (err & 3).any()
    if chance == 'replay':
        pos = to_np(env.tape_pos)
        for i in range(S):
            assert pos[i] == len(tapes[i])
    print('%s deep fixture on the device (%s): %d slots, %d episodes, %d records' % (
        game, chance, S, int(fx['slot_episodes'][slots].sum()), checked))


class FacadeAdapter:
    def __init__(self, env):
        self.env = env; self.state = None
    def reset(self):
        self.state, _ = self.env.reset()
    def step(self, a):
        self.state, _ = self.env.step(a)
    def obs(self, seat):
        st = self.state if seat < 0 else self.env.get_state(seat)
        self._last = st
        return np.asarray(st['obs']).reshape(-1)
    def legal_mask(self):
        m = np.zeros(self.env.num_actions, np.uint8); m[list(self._last['legal_actions'].keys())] = 1
        return m
    def is_over(self):
        return self.env.is_over()
    def player(self):
        return self.env.get_player_id()
    def payoffs(self):
        return self.env.get_payoffs()


@pytest.mark.parametrize('game', FIX_GAMES)
def test_make_facade_reproduces_reference_from_seed(game):
    """rlcard_b200.make(env, {'seed': s}) == rlcard.make(env, {'seed': s}): np.random.RandomState runs on
    the device (MT19937 + numpy legacy bounded draws), no tape."""
    fx = load_fixture(game)
    game = fixture_game(game)
    for slot in range(min(3, len(fx['slot_seed']))):
        env = rlcard_b200.make(game, {'seed': int(fx['slot_seed'][slot])})
        check_slot(fx, slot, FacadeAdapter(env), game + ' (facade)')


@pytest.mark.parametrize('obs_dtype', [torch.uint8, torch.float32])
@pytest.mark.parametrize('game', GAMES)
def test_throughput_rollout_equals_oracle(game, obs_dtype):
    """Throughput mode (Philox chance + fused random policy + auto reset) vs the CPU twin, bit-exact on the
    whole trajectory, across two consecutive launches (state carried in HBM between them)."""
    if rlcard_b200.game_info(game).obs_native_dtype == 1 and obs_dtype == torch.uint8:
        pytest.skip('fractional obs')
    long_game = game in ('doudizhu', 'scout', 'uno')
    n, T, seed, base = 1000, (25 if long_game else 40), 987654321, 77          # ragged: n not a multiple of 32
    env = rlcard_b200.VecEnv(game, n, seed=seed, env_id_base=base, obs_dtype=obs_dtype)
    orc = oracle.OracleVec(game, n, seed, env0=base)
    env.reset()
    episodes = 0
    for launch in range(6 if long_game else 2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        episodes += int(ref['done'].sum())
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            got, want = to_np(tr[k]), ref[k]
            if k == 'mask' and env.mask_bitpacked:                     # compare bit-packed (27 472 ids per row)
                want = np.packbits(want, axis=-1, bitorder='little')
                got = got.view(np.uint8)[..., :want.shape[-1]]
            if k == 'obs':                                   # row stride may be padded beyond the widest seat
                D = want.shape[-1]
                assert not got[..., D:].any()
                got = got[..., :D]
                want = want.astype(got.dtype)
                assert np.array_equal(want.astype(np.float32), ref[k])
            assert np.array_equal(got, want.astype(got.dtype)), '%s launch %d %s' % (game, launch, k)
    env.check_errors()
    assert episodes > 0


def test_uno_long_rollout_reshuffles_equal_oracle():
    """UNO throughput mode long enough for replace_deck (played pile back into the draw pile) to happen many
    times: the multiset-pile kernel and the oracle's list piles must agree on every action / player / done /
    payoff and on obs + mask at the end."""
    if not built('uno'):
        pytest.skip('uno not built')
    n, T, seed = 192, 700, 55
    env = rlcard_b200.VecEnv('uno', n, seed=seed)
    env.reset()
    tr = env.rollout_random(T)
    ref = oracle.OracleVec('uno', n, seed).rollout(T, nthreads=4)
    for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
        np.testing.assert_array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64), err_msg=k)
    env.check_errors()
    assert int(ref['action'][ref['action'] == 60].size) > 1000          # plenty of draws


def test_host_rollout_equals_device_rollout():
    """VecEnv.rollout_random_host (chunked launches + overlapped D2H into pinned memory, state owned by the
    host) delivers the same trajectory as one device-side rollout."""
    n, T, seed = 4096, 50, 77
    a = rlcard_b200.VecEnv('leduc-holdem', n, seed=seed)
    b = rlcard_b200.VecEnv('leduc-holdem', n, seed=seed)
    a.reset(); b.reset()
    ref = a.rollout_random(T)
    h_state = torch.empty(b.state.shape, dtype=b.state.dtype).pin_memory()
    h_state.copy_(b.state)
    torch.cuda.synchronize()
    out = b.rollout_random_host(T, b.alloc_host_trajectory(T), chunk=16, host_state=h_state)
    for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
        assert torch.equal(out[k], ref[k].cpu()), k
    assert torch.equal(h_state, a.state.cpu())


@pytest.mark.parametrize('game', ['leduc-holdem', 'limit-holdem', 'uno', 'doudizhu', 'scout'])
@pytest.mark.parametrize('obs_dtype', [torch.uint8, torch.float32])
def test_compact_host_rollout_expands_to_dense_rows(game, obs_dtype):
    """rollout_random_host(compact=True): the records that cross PCIe (4 / 12 / 28 / 132 / 80 bytes per env-step) expand on the
    host (rlcard_b200.compact.expand) into exactly the dense trajectory a device-side rollout of the same envs produces --
    for DouDizhu including the 27 472-bit legal mask, which is not sent but recomputed from the row and the action table."""
    from rlcard_b200 import compact
    if rlcard_b200.game_info(game).obs_native_dtype == 1 and obs_dtype == torch.uint8:
        pytest.skip('fractional obs')
    n, T, seed = (515, 40, 99) if game == 'doudizhu' else (4099, 70, 99)     # ragged batch, more steps than one chunk
    a = rlcard_b200.VecEnv(game, n, seed=seed, obs_dtype=obs_dtype)
    b = rlcard_b200.VecEnv(game, n, seed=seed, obs_dtype=obs_dtype)
    a.reset(); b.reset()
    ref = a.rollout_random(T)
    packed = b.rollout_random_host(T, b.alloc_host_compact(T), chunk=16, compact=True)
    assert packed.shape == (T, n, b.compact_words()) and packed.numel() * 4 * 4 < ref['obs'].numel()
    got = compact.expand(game, packed.numpy())
    for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
        want = to_np(ref[k])
        if k == 'mask' and game == 'doudizhu':
            want = want.view(np.uint32)
        assert got[k].shape == want.shape, (game, k, got[k].shape, want.shape)
        assert np.array_equal(got[k].astype(want.dtype), want), (game, k)
    assert int(to_np(ref['done']).sum()) > 0
    assert torch.equal(a.state, b.state)
    with pytest.raises(rlcard_b200.RlcError):
        rlcard_b200.VecEnv('blackjack', 64, seed=1).alloc_host_compact(4)


@pytest.mark.parametrize('game', GAMES)
def test_step_api_equals_fused_rollout(game):
    """Env.step kernel == fused rollout kernel: feeding the rollout's actions through step() (auto reset)
    reproduces obs / mask / player / done / payoffs."""
    n, T, seed = 512, 30, 4242
    a = rlcard_b200.VecEnv(game, n, seed=seed, obs_dtype=torch.float32)
    b = rlcard_b200.VecEnv(game, n, seed=seed, obs_dtype=torch.float32)
    a.reset()
    tr = a.rollout_random(T)
    obs, mask, cur = b.reset()
    for t in range(T):
        assert torch.equal(obs, tr['obs'][t]) and torch.equal(mask, tr['mask'][t]) and torch.equal(cur, tr['player'][t]), (game, t)
        obs, mask, cur, done, pay = b.step(tr['action'][t])
        assert torch.equal(done, tr['done'][t]) and torch.equal(pay, tr['payoffs'][t]), (game, t)
    assert torch.equal(a.state, b.state)


@pytest.mark.parametrize('game', GAMES)
def test_full_size_properties(game):
    """BASELINE.json sizes: properties that do not need the oracle."""
    n = {'leduc-holdem': 65536, 'limit-holdem': 16384, 'uno': 16384, 'blackjack': 65536, 'no-limit-holdem': 16384}.get(game, 8192)
    T = 64
    env = rlcard_b200.VecEnv(game, n, seed=7)
    env.reset()
    tr = env.rollout_random(T)
    env.check_errors()
    mask, act, done, pay = tr['mask'], tr['action'].long(), tr['done'].bool(), tr['payoffs']
    if env.mask_bitpacked:
        word = mask.gather(2, (act >> 5).unsqueeze(-1)).squeeze(-1)
        assert bool((((word >> (act & 31).int()) & 1) == 1).all())         # every action taken was legal
        assert bool(((mask != 0).sum(-1) >= 1).all())
    else:
        assert bool((mask.gather(2, act.unsqueeze(-1)) == 1).all())
        assert bool((mask.sum(-1) >= 1).all())
    assert bool((pay[~done] == 0).all())
    if game in ('leduc-holdem', 'limit-holdem', 'no-limit-holdem'):
        assert bool((pay.sum(-1) == 0).all())                              # zero-sum
        assert bool(((pay * 2) == (pay * 2).round()).all())                # multiples of 0.5
    assert int(done.sum()) > (n if game in ('blackjack', 'leduc-holdem', 'limit-holdem', 'no-limit-holdem') else 0)   # episodes finish and restart
    # shard invariance: envs [n/2, n) of this run == a second VecEnv with env_id_base n/2
    env2 = rlcard_b200.VecEnv(game, n // 2, seed=7, env_id_base=n // 2)
    env2.reset()
    tr2 = env2.rollout_random(T)
    for k in ('obs', 'mask', 'action', 'done', 'payoffs'):
        assert torch.equal(tr[k][:, n // 2:], tr2[k]), k


@pytest.mark.parametrize('game', GAMES)
def test_full_size_rollout_equals_oracle(game):
    """BASELINE.json's sizes (65 536 / 16 384 / 8 192 envs per GPU), T = 16: every trajectory tensor of the fused
    rollout equals the CPU twin, after a first launch that scatters the envs over their episodes."""
    n = {'leduc-holdem': 65536, 'limit-holdem': 16384, 'uno': 16384, 'blackjack': 65536, 'no-limit-holdem': 16384}.get(game, 8192)
    T, seed = (8 if game == 'doudizhu' else 16), 20261019         # DouDizhu: the oracle's dense mask is 27 472 bytes per row
    env = rlcard_b200.VecEnv(game, n, seed=seed)
    orc = oracle.OracleVec(game, n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=16)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            got, want = to_np(tr[k]), ref[k]
            if k == 'mask' and env.mask_bitpacked:
                want = np.packbits(want, axis=-1, bitorder='little')
                got = got.view(np.uint8)[..., :want.shape[-1]]
            if k == 'obs':
                assert not got[..., want.shape[-1]:].any()
                got = got[..., :want.shape[-1]]
            assert np.array_equal(got, want.astype(got.dtype)) and np.array_equal(want.astype(got.dtype).astype(want.dtype), want), \
                '%s full size launch %d %s' % (game, launch, k)
        del tr, ref
    env.check_errors()
    print('%s: %d envs x %d steps x 2 launches == oracle' % (game, n, T))


@pytest.mark.parametrize('game', GAMES)
def test_long_horizon_rollout_equals_oracle(game):
    """1024 env-steps per env over four launches (hundreds of episodes per env for the short games, UNO reshuffles,
    Philox step counters far past the first blocks): actions, players, dones and payoffs of every step and the obs /
    mask rows of the last launch equal the CPU twin."""
    n, T, launches, seed, base = 512, 256, 4, 20261018, 123457
    env = rlcard_b200.VecEnv(game, n, seed=seed, env_id_base=base)
    orc = oracle.OracleVec(game, n, seed, env0=base)
    env.reset()
    for launch in range(launches):
        last = launch == launches - 1
        tr = env.rollout_random(T)
        ref = orc.rollout(T, want_obs=last, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs') + (('mask', 'obs') if last else ()):
            got, want = to_np(tr[k]), ref[k]
            if k == 'mask' and env.mask_bitpacked:
                want = np.packbits(want, axis=-1, bitorder='little')
                got = got.view(np.uint8)[..., :want.shape[-1]]
            if k == 'obs':
                got = got[..., :want.shape[-1]]
            assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), '%s launch %d %s' % (game, launch, k)
    env.check_errors()


@pytest.mark.parametrize('game', GAMES)
def test_rollout_does_not_depend_on_stale_shared_memory(game):
    """Two identical rollouts with other kernels in between (which leave their own bytes in shared memory) must agree:
    tiles that rows are OR-ed into are zeroed by the kernel itself, never assumed clean."""
    n, T, seed = 4096, 32, 31337
    a = rlcard_b200.VecEnv(game, n, seed=seed)
    a.reset()
    ta = a.rollout_random(T)
    for other in GAMES:                                            # dirty every SM's shared memory with other layouts
        if other != game:
            o = rlcard_b200.VecEnv(other, 8192, seed=1)
            o.reset()
            o.rollout_random(4)
    b = rlcard_b200.VecEnv(game, n, seed=seed)
    b.reset()
    tb = b.rollout_random(T)
    for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
        assert torch.equal(ta[k], tb[k]), (game, k)
    assert torch.equal(a.state, b.state)


def test_illegal_action_fallback_and_noop():
    env = rlcard_b200.VecEnv('leduc-holdem', 64, seed=3, auto_reset=False)
    obs, mask, cur = env.reset()
    before = env.state.clone()
    env.step(torch.full((64,), -1, dtype=torch.int32, device='cuda'))     # no-op
    assert torch.equal(before, env.state)
    env.step(torch.full((64,), 3, dtype=torch.int32, device='cuda'))      # 'check' is illegal for the small blind
    assert bool((env.err & 4).all())                                       # -> fold fallback (envs/leducholdem.py:90-96)
    assert bool(env.done.all())


@pytest.mark.parametrize('game', GAMES)
def test_partial_streams_and_unaligned_batches(game):
    """Kernel variants: rollouts that request only some trajectory streams (generic instantiation with null checks),
    and batch sizes whose obs rows are not 16-byte aligned (byte-wise flush; for Leduc the register engine instead of
    the tabulated one) must produce the same envs as the full, aligned launch (Philox streams are keyed by env id)."""
    T, seed = 24, 99
    n_full = 256
    full = rlcard_b200.VecEnv(game, n_full, seed=seed)
    full.reset()
    ref = full.rollout_random(T)
    # (1) only action / done / payoffs requested
    lean = rlcard_b200.VecEnv(game, n_full, seed=seed)
    lean.reset()
    out = lean.rollout_random(T, out=lean.alloc_trajectory(T, obs=False, mask=False))
    for k in ('action', 'player', 'done', 'payoffs'):
        assert torch.equal(out[k], ref[k]), (game, k)
    assert torch.equal(lean.state, full.state)
    # (2) odd batch size: envs [0, 131) of the same id range
    n_odd = 131
    odd = rlcard_b200.VecEnv(game, n_odd, seed=seed)
    odd.reset()
    out = odd.rollout_random(T)
    for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
        assert torch.equal(out[k], ref[k][:, :n_odd]), (game, k)
    odd.check_errors()


@pytest.mark.timeout(300)
def test_limit_warp_specialised_rollout_equals_oracle(monkeypatch):
    """RLC_LIMIT_WS=1: the four-role Limit rollout (ENV / EMIT / DEAL / JUDGE warps coupled by shared-memory rings,
    tu_limit.cu) gives the oracle's trajectory, over two launches (episodes under way at the launch boundary)."""
    monkeypatch.setenv('RLC_LIMIT_WS', '1')
    n, T, seed = 2048, 48, 606
    env = rlcard_b200.VecEnv('limit-holdem', n, seed=seed)
    orc = oracle.OracleVec('limit-holdem', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (launch, k)
    env.check_errors()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('epw', [16, 32])
def test_uno_env_emit_rollout_equals_oracle(epw, monkeypatch):
    """RLC_UNO_EE=16|32: the two-warp UNO rollout (ENV transition warp + EMIT row warp, tu_uno.cu) gives the oracle's
    trajectory over two launches (episodes under way at the launch boundary)."""
    monkeypatch.setenv('RLC_UNO_EE', str(epw))
    n, T, seed = 1024, 60, 707
    env = rlcard_b200.VecEnv('uno', n, seed=seed)
    orc = oracle.OracleVec('uno', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (launch, k)
    env.check_errors()


@pytest.mark.timeout(120)
@pytest.mark.parametrize('T', [60, 5])
@pytest.mark.parametrize('epw', [16, 32])
def test_uno_chunked_env_emit_rollout_equals_oracle(epw, T, monkeypatch):
    """RLC_UNO_PIPE=16|32: the ENV / EMIT UNO rollout with chunked hand-over (named barriers once per 8 env-steps, tu_uno.cu) gives
    the oracle's trajectory over two launches, for windows longer and shorter than a chunk."""
    monkeypatch.setenv('RLC_UNO_PIPE', str(epw))
    n, seed = 1024, 709
    env = rlcard_b200.VecEnv('uno', n, seed=seed)
    orc = oracle.OracleVec('uno', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (launch, k)
    env.check_errors()


@pytest.mark.parametrize('bulk', [0, 1, 3])
def test_doudizhu_bulk_row_variants_equal_oracle(bulk, monkeypatch):
    """RLC_WROLLOUT_BULK: the DouDizhu rollout with its mask row (1) or mask + obs rows (3) leaving shared memory as
    cp.async.bulk copies must equal the LDS + STG variant (0) and the oracle."""
    monkeypatch.setenv('RLC_WROLLOUT_BULK', str(bulk))
    n, T, seed = 600, 30, 808
    env = rlcard_b200.VecEnv('doudizhu', n, seed=seed)
    orc = oracle.OracleVec('doudizhu', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            got, want = to_np(tr[k]), ref[k]
            if k == 'mask':
                want = np.packbits(want, axis=-1, bitorder='little')
                assert not got.view(np.uint8)[..., want.shape[-1]:].any()            # pad bits / pad word stay zero
                got = got.view(np.uint8)[..., :want.shape[-1]]
            if k == 'obs':
                got = got[..., :want.shape[-1]]
            assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), (bulk, launch, k)
    env.check_errors()


@pytest.mark.parametrize('n', [203, 512])
@pytest.mark.parametrize('lpe', [32, 16, 8])
def test_scout_lanes_per_env_variants_equal_oracle(lpe, n, monkeypatch):
    """RLC_WROLLOUT_LPE: the Scout throughput rollout with a warp (32), a half-warp (16) or a quarter-warp (8, default) per
    env (kernels_warp.cuh k_wrollout_multi) gives the oracle's trajectory and final state -- on a ragged batch (ghost
    groups in the last warp) and an aligned one, over three launches (episodes of ~146 steps span them)."""
    monkeypatch.setenv('RLC_WROLLOUT_LPE', str(lpe))
    T, seed = 64, 909
    env = rlcard_b200.VecEnv('scout', n, seed=seed)
    orc = oracle.OracleVec('scout', n, seed)
    env.reset()
    for launch in range(3):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (lpe, launch, k)
    env.check_errors()


@pytest.mark.parametrize('block', [64, 128, 256])
@pytest.mark.parametrize('fsm', ['0', '1', '2'])
def test_limit_tabulated_rollout_variants_equal_oracle(fsm, block, monkeypatch):
    """RLC_LIMIT_FSM: the generic register engine (0), the tabulated automaton with deal rings on one warp per group (1)
    and split over two warps per group (2, default) give the oracle's trajectory, on a ragged batch, over three launches
    (state words carried across launches, deals keyed by the episode ordinal), for every block size of the sweep."""
    monkeypatch.setenv('RLC_LIMIT_FSM', fsm)
    monkeypatch.setenv('RLC_LIMIT_BLOCK', str(block))
    n, T, seed = 1000, 40, 616
    env = rlcard_b200.VecEnv('limit-holdem', n, seed=seed)
    orc = oracle.OracleVec('limit-holdem', n, seed)
    env.reset()
    for launch in range(3):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (fsm, block, launch, k)
    env.check_errors()


@pytest.mark.parametrize('T', [40, 7, 129])
@pytest.mark.parametrize('pipe', ['32', '22', '42', '31'])
def test_limit_pipelined_rollout_equals_oracle(pipe, T, monkeypatch):
    """RLC_LIMIT_PIPE=<EMIT warps><DEAL warps>: the tabulated Limit rollout as a pipeline of role warps per group (ENV automaton ->
    16-step record chunks -> EMIT warps; DEAL warps fill tagged deal rings ahead) gives the oracle's trajectory on a ragged batch,
    over three launches (state carried), for window lengths below, at and beyond one chunk."""
    monkeypatch.setenv('RLC_LIMIT_PIPE', pipe)
    n, seed = 1000, 636
    env = rlcard_b200.VecEnv('limit-holdem', n, seed=seed)
    orc = oracle.OracleVec('limit-holdem', n, seed)
    env.reset()
    for launch in range(3):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (pipe, launch, k)
    env.check_errors()


def test_limit_tabulated_rollout_continues_step_api_state():
    """The tabulated Limit rollout maps the packed state words onto its automaton and back: envs advanced by rlc_step
    (mid-episode, and finished episodes left without auto reset) continue exactly like the oracle."""
    n, seed = 777, 626
    env = rlcard_b200.VecEnv('limit-holdem', n, seed=seed)
    orc = oracle.OracleVec('limit-holdem', n, seed)
    env.reset()
    tr = env.rollout_random(7)
    ref = orc.rollout(7, nthreads=8)
    assert np.array_equal(to_np(tr['action']), ref['action'])
    tr = env.rollout_random(33)
    ref = orc.rollout(33, nthreads=8)
    for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
        assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), k
    env.check_errors()


@pytest.mark.parametrize('block', [64, 128, 512, 1024])
def test_leduc_block_size_variants_equal_oracle(block, monkeypatch):
    """RLC_LEDUC_BLOCK: the tabulated Leduc rollout is launched with one SMSP-balanced block per SM at large batches (512 threads
    at 65 536 envs) and with 64-thread blocks at small ones; every block size gives the oracle's trajectory on a ragged batch."""
    monkeypatch.setenv('RLC_LEDUC_BLOCK', str(block))
    n, T, seed = 5003, 40, 717
    env = rlcard_b200.VecEnv('leduc-holdem', n, seed=seed)
    orc = oracle.OracleVec('leduc-holdem', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (block, launch, k)
    env.check_errors()


@pytest.mark.parametrize('block', [64, 128, 256, 448, 512])
def test_scout_block_size_variants_equal_oracle(block, monkeypatch):
    """RLC_WROLLOUT_BLOCK: block sizes of the four-envs-per-warp Scout rollout (448 = one 14-warp block per SM at 8 192 envs)."""
    monkeypatch.setenv('RLC_WROLLOUT_BLOCK', str(block))
    n, T, seed = 333, 48, 727
    env = rlcard_b200.VecEnv('scout', n, seed=seed)
    orc = oracle.OracleVec('scout', n, seed)
    env.reset()
    for launch in range(2):
        tr = env.rollout_random(T)
        ref = orc.rollout(T, nthreads=8)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            assert np.array_equal(to_np(tr[k]).astype(np.float64), ref[k].astype(np.float64)), (block, launch, k)
    env.check_errors()


THREAD_GAMES = [g for g in GAMES if g not in ('doudizhu', 'scout')]


@pytest.mark.parametrize('epw', [32, 16, 8])
@pytest.mark.parametrize('game', THREAD_GAMES)
def test_rollout_envs_per_warp_variants_equal_oracle(game, epw, monkeypatch):
    """The fused rollout of the thread-per-env games runs 32, 16 or 8 envs per warp depending on the batch size
    (kernels.cuh rollout_envs_per_warp; RLC_ROLLOUT_EPW forces one).  Every variant must give the oracle's trajectory,
    on a ragged batch (last warp partly filled) and on an aligned one (compile-time tile flush)."""
    monkeypatch.setenv('RLC_ROLLOUT_EPW', str(epw))
    T, seed = 30, 424242
    for n in (203, 512):
        env = rlcard_b200.VecEnv(game, n, seed=seed)
        env.reset()
        tr = env.rollout_random(T)
        ref = oracle.OracleVec(game, n, seed).rollout(T, nthreads=4)
        for k in ('action', 'player', 'done', 'payoffs', 'mask', 'obs'):
            got, want = to_np(tr[k]), ref[k]
            if k == 'obs':
                got = got[..., :want.shape[-1]]
            assert np.array_equal(got.astype(np.float64), want.astype(np.float64)), (game, epw, n, k)
        env.check_errors()


def test_reset_mask_and_tape_errors():
    """rlc_reset with a partial mask only deals the selected envs; an exhausted replay tape raises the error flag."""
    env = rlcard_b200.VecEnv('leduc-holdem', 64, seed=1, auto_reset=False)
    env.reset()
    before = env.state.clone()
    m = torch.zeros(64, dtype=torch.uint8, device='cuda'); m[::2] = 1
    env.reset(m)
    changed = (env.state != before).any(0).cpu().numpy()
    assert not changed[1::2].any() and changed[::2].all()          # episode counter moved only where asked
    rep = rlcard_b200.VecEnv('leduc-holdem', 4, mode='replay', auto_reset=False)
    rep.set_tape(np.zeros((4, 2), np.uint8))                          # a Leduc deal needs 6 draws
    rep.reset()
    assert bool((rep.err & 1).all())
    with pytest.raises(rlcard_b200.RlcError):
        rep.check_errors()


def test_illegal_actions_in_every_game_are_flagged():
    """Out-of-range ids never corrupt the state: the per-env err flag 4 is raised and the env stays consistent
    (the reference raises or substitutes, envs/leducholdem.py:90-96, envs/uno.py:39-45, games/scout/round.py:92)."""
    for game in GAMES:
        if game == 'blackjack':
            continue                                                  # both ids are always legal
        env = rlcard_b200.VecEnv(game, 32, seed=5, auto_reset=False)
        env.reset()
        env.step(torch.full((32,), env.num_actions + 7, dtype=torch.int32, device='cuda'))
        assert bool((env.err & 4).all()), game
        obs, mask, cur, done, pay = env.get_state(None)
        m = mask if not env.mask_bitpacked else (mask != 0)
        assert bool(((m != 0).sum(-1) >= 1)[~done.bool()].all()), game


@pytest.mark.parametrize('game', ['leduc-holdem', 'uno', 'doudizhu'])
def test_checkpoint_resume_is_bit_exact(game):
    """state_dict / load_state_dict: a VecEnv restored into a fresh object continues with the same trajectory."""
    n, T = 300, 20
    a = rlcard_b200.VecEnv(game, n, seed=8)
    a.reset()
    a.rollout_random(T)
    ck = a.state_dict()
    want = a.rollout_random(T)
    b = rlcard_b200.VecEnv(game, n, seed=12345)                       # seed comes from the checkpoint
    b.load_state_dict(ck)
    assert torch.equal(b.obs, want['obs'][0]) and torch.equal(b.cur_player, want['player'][0])
    got = b.rollout_random(T)
    for k in ('obs', 'mask', 'action', 'player', 'done', 'payoffs'):
        assert torch.equal(got[k], want[k]), (game, k)
