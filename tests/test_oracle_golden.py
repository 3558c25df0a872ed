"""Pins the CPU oracle (oracle/*.c) against fixtures recorded from the live reference."""
import numpy as np
import pytest

import oracle
from replay_util import ALL_FIXTURES, check_slot, fixture_game, have_fixture, load_fixture, slot_records, slot_tape

GAMES = [g for g in ALL_FIXTURES if have_fixture(g)]


@pytest.mark.parametrize('game', GAMES)
def test_oracle_replays_reference_tape(game):
    fx = load_fixture(game)
    total = 0
    for slot in range(len(fx['slot_seed'])):
        env = oracle.OracleEnv(fixture_game(game))
        tape = slot_tape(fx, slot)
        env.set_tape(tape)
        total += check_slot(fx, slot, env, game)
        assert env.tape_err() == 0
        assert env.tape_pos() == len(tape)
    assert total == len(fx['rec_slot'])


@pytest.mark.parametrize('game', GAMES)
def test_oracle_mt19937_reproduces_reference_from_seed(game):
    """seed -> sha512 -> MT19937 -> numpy-legacy shuffle/randint/choice: the draws the oracle makes
    from the seed alone equal the draws the reference's RandomState made."""
    fx = load_fixture(game)
    for slot in range(len(fx['slot_seed'])):
        env = oracle.OracleEnv(fixture_game(game))
        env.seed(int(fx['slot_seed'][slot]))
        env.record()
        check_slot(fx, slot, env, game + ' (mt)')
        np.testing.assert_array_equal(env.recorded(), slot_tape(fx, slot))


def test_philox_known_answers():
    L = oracle.lib()
    out = np.zeros(4, np.uint32)
    ctr = np.zeros(4, np.uint32)
    L.orc_philox4x32_10(ctr.ctypes.data, 0, 0, out.ctypes.data)
    assert [hex(v) for v in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    ctr[:] = 0xffffffff
    L.orc_philox4x32_10(ctr.ctypes.data, 0xffffffff, 0xffffffff, out.ctypes.data)
    assert [hex(v) for v in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    ctr[:] = [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]
    L.orc_philox4x32_10(ctr.ctypes.data, 0xa4093822, 0x299f31d0, out.ctypes.data)
    assert [hex(v) for v in out] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_seed_words():
    """rlcard/utils/seeding.py:91-113 word split: 1-2 little-endian uint32 words of sha512(str(seed))[:8]."""
    import hashlib, struct
    for seed in (0, 1, 42, 12941, 2 ** 40 + 7):
        w = oracle.seed_words(seed)
        lo, hi = struct.unpack('2I', hashlib.sha512(str(seed).encode()).digest()[:8])
        assert w == ([lo, hi] if hi else [lo])


def test_doudizhu_type_prefilter_is_sound():
    """The device legal-set generator skips whole action types the hand cannot play (game_doudizhu.cuh ddz_type_ok ==
    doudizhu_table.type_feasible).  Sound iff every table row passes the requirement of its own type: a hand that
    contains the row then passes it too (the requirement is monotone in the rank counts).  Also pins the blob layout
    the library parses (need pairs: nibble-wise minimum | type bits per mask word and per batch of 32 words)."""
    import struct
    import numpy as np
    from rlcard_b200 import doudizhu_table as dt
    tab = dt.load()
    for i, x in enumerate(tab['counts']):
        if i != dt.PASS_ID:
            assert dt.type_feasible(dt.unpack_counts(x), int(tab['type'][i])), (i, dt.TYPE_NAMES[tab['type'][i]])
    for t, name in enumerate(dt.TYPE_NAMES):                       # the empty hand plays nothing
        assert not dt.type_feasible([0] * 15, t), name
    blob = dt.build_blob()
    magic, n_actions, n_words, n_types, off_rows, off_need, off_type, off_weight, off_tw, total = struct.unpack_from('<4sIII6Q', blob)
    assert magic == b'DDZ2' and n_actions == 27472 and n_words == 896 and n_types == 38 and total == len(blob)
    need = np.frombuffer(blob, np.uint64, 2 * 896, off_need).reshape(896, 2)
    types = tab['type'].astype(np.int64)
    for j in (0, 1, 400, 858):
        want = 0
        for i in range(32 * j, min(32 * j + 32, dt.NUM_ACTIONS)):
            want |= 1 << int(types[i])
        assert int(need[j, 1]) == want
    assert int(need[864 + 26, 1]) & (1 << dt.T_PASS) and not int(need[864, 1]) & (1 << dt.T_PASS)


# ---- deep fixtures: >= 1e4 / 1e3 / 500 reference episodes per game + stratified / crafted slots -------------------
from replay_util import DEEP_GAMES, check_slot_deep, load_deep  # noqa: E402


@pytest.mark.parametrize('game', DEEP_GAMES)
def test_oracle_replays_deep_reference_fixture(game):
    fx = load_deep(game)
    total = 0
    for slot in range(len(fx['slot_seed'])):
        env = oracle.OracleEnv(game)
        tape = fx['tape'][fx['tape_off'][slot]:fx['tape_off'][slot + 1]]
        env.set_tape(tape)
        total += check_slot_deep(fx, slot, env, game + ' (deep)')
        assert env.tape_err() == 0
        assert env.tape_pos() == len(tape)
    assert total == len(fx['rec_kind'])
    print('%s deep fixture: %d slots, %d episodes, %d records (%d env-steps) replayed bit-exact' % (
        game, len(fx['slot_seed']), int(fx['slot_episodes'].sum()), total, int((fx['rec_kind'] == 1).sum())))


@pytest.mark.parametrize('game', DEEP_GAMES)
def test_oracle_mt19937_regenerates_deep_tapes(game):
    """every 5th seeded slot: the draws made from the seed alone equal the recorded tape"""
    fx = load_deep(game)
    for slot in range(0, len(fx['slot_seed']), 5):
        if fx['slot_seed'][slot] < 0:
            continue
        env = oracle.OracleEnv(game)
        env.seed(int(fx['slot_seed'][slot]))
        env.record()
        check_slot_deep(fx, slot, env, game + ' (deep, mt)')
        np.testing.assert_array_equal(env.recorded(), fx['tape'][fx['tape_off'][slot]:fx['tape_off'][slot + 1]])
