"""Known-answer tests harvested from the reference's own unit tests and scored by the live reference
(tests/golden/make_kats.py -> tests/golden/kats.npz): the CPU oracle on every vector (not gpu), and the CUDA
judger operators through the C ABI (gpu)."""
import os

import numpy as np
import pytest

import oracle

KATS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kats.npz')


@pytest.fixture(scope='module')
def kats():
    z = np.load(KATS)
    return {k: z[k] for k in z.files}


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_holdem_compare_hands(kats):
    L = oracle.lib()
    cards, nps, want = kats['holdem_cards'], kats['holdem_np'], kats['holdem_winners']
    assert int((kats['holdem_src'] == 0).sum()) >= 100        # the reference's own test vectors are in there
    for i in range(len(cards)):
        P = int(nps[i])
        c = np.ascontiguousarray(cards[i, :P])
        win = np.zeros(P, np.uint8)
        L.orc_holdem_winners(c.ctypes.data, P, win.ctypes.data)
        assert win.tolist() == want[i, :P].tolist(), (i, cards[i, :P].tolist())


def test_oracle_doudizhu_playable_sets(kats):
    L = oracle.lib()
    oracle.need_tables('doudizhu')
    hands, tg, legal = kats['ddz_hand'], kats['ddz_target'], kats['ddz_legal']
    masks = np.zeros((len(hands), 27472), np.uint8)
    for i in range(len(hands)):
        h = np.ascontiguousarray(hands[i])
        L.orc_doudizhu_legal_for(h.ctypes.data, int(tg[i]), masks[i].ctypes.data)
    want = np.unpackbits(legal, axis=1, bitorder='little')[:, :27472]
    bad = np.nonzero((masks != want).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], hands[bad[:5]], tg[bad[:5]])
    for case, action, expect in kats['ddz_assert']:            # the reference test file's in / not-in assertions
        assert masks[case, action] == expect
    assert int(kats['ddz_n_test_hands']) >= 6 and len(kats['ddz_assert']) >= 170


def test_host_doudizhu_legal_mask_matches_reference(kats):
    """rlcard_b200.compact.doudizhu_legal_mask (the host side of DouDizhu's compact wire format recomputes the legal set from
    the hand and the action to beat) on the reference's 1 607 playable-set vectors."""
    from rlcard_b200 import compact, doudizhu_table as T
    tab = T.load()
    hands, tg, legal = kats['ddz_hand'], kats['ddz_target'], kats['ddz_legal']
    for i in range(len(hands)):
        target = [0] * 15 if tg[i] < 0 else T.unpack_counts(tab['counts'][tg[i]])
        got = compact.doudizhu_legal_mask(hands[i], target).view(np.uint8)[:legal.shape[1]]
        assert np.array_equal(got, legal[i]), (i, hands[i], tg[i])


def test_oracle_leduc_judger(kats):
    L = oracle.lib()
    out = np.zeros(2, np.float64)
    for c, want in zip(kats['leduc_case'], kats['leduc_payoffs']):
        L.orc_leduc_judge(*[int(x) for x in c], out.ctypes.data)
        assert (out * 2).tolist() == want.tolist(), c          # judge_game is in chips, the env payoff in big blinds


def test_oracle_uno_encoders(kats):
    L = oracle.lib()
    o = np.zeros(240, np.float32)
    for h, t, want in zip(kats['uno_hand'], kats['uno_target'], kats['uno_obs']):
        codes = np.ascontiguousarray(h[h != 255])
        L.orc_uno_encode(codes.ctypes.data, len(codes), int(t), o.ctypes.data)
        assert np.array_equal(o.astype(np.uint8), want), h


# ------------------------------------------------------------------ CUDA operators through the C ABI
@pytest.mark.gpu
def test_cuda_holdem_compare_hands(kats):
    import torch
    from rlcard_b200 import judgers
    cards, nps, want, dup = kats['holdem_cards'], kats['holdem_np'], kats['holdem_winners'], kats['holdem_dup']
    checked = 0
    for P in (2, 3, 4):
        idx = np.nonzero((nps == P) & (dup == 0))[0]           # a physical card twice cannot be dealt
        if len(idx) == 0:
            continue
        got = judgers.compare_hands(torch.from_numpy(cards[idx, :P]).cuda()).cpu().numpy()
        assert np.array_equal(got, want[idx, :P]), np.nonzero((got != want[idx, :P]).any(1))[0][:5]
        checked += len(idx)
    assert checked > 29000


@pytest.mark.gpu
def test_cuda_holdem_table_evaluator_random_hands():
    """400 000 random showdowns (2 seats, 5 shared board cards) through rlc_judge_holdem: the table evaluator (card64 / t5 in shared
    memory) must agree with the branch-free evaluator on every hand (the kernel answers 255 where they differ) and with the
    oracle's evaluator on the winners."""
    import torch
    import oracle
    from rlcard_b200 import judgers
    rng = np.random.default_rng(20261019)
    n = 400_000
    deck = np.argsort(rng.random((n, 52)), axis=1)[:, :9].astype(np.uint8)          # nine distinct cards per case
    cards = np.empty((n, 2, 7), np.uint8)
    cards[:, 0, :2], cards[:, 1, :2] = deck[:, 0:2], deck[:, 2:4]
    cards[:, :, 2:] = deck[:, None, 4:9]
    got = judgers.compare_hands(torch.from_numpy(cards).cuda()).cpu().numpy()
    assert not (got == 255).any(), np.nonzero((got == 255).any(1))[0][:5]
    L = oracle.lib()
    want = np.zeros((n, 2), np.uint8)
    flat = np.ascontiguousarray(cards.reshape(n, 14))
    for i in range(0, n, 97):                                                       # the oracle on a 1 % sample
        L.orc_holdem_winners(flat[i].ctypes.data, 2, want[i].ctypes.data)
        assert np.array_equal(got[i], want[i]), (i, cards[i])
    ties = (got.sum(1) == 2).mean()
    assert 0.005 < ties < 0.08                                                       # split pots exist and are rare


@pytest.mark.gpu
def test_cuda_doudizhu_playable_sets(kats):
    import torch
    from rlcard_b200 import judgers
    got = judgers.doudizhu_playable(torch.from_numpy(kats['ddz_hand']).cuda(), torch.from_numpy(kats['ddz_target']).cuda())
    got = got.cpu().numpy().view(np.uint8)[:, :kats['ddz_legal'].shape[1]]
    assert np.array_equal(got, kats['ddz_legal'])


@pytest.mark.gpu
def test_cuda_leduc_judger(kats):
    import torch
    from rlcard_b200 import judgers
    got = judgers.judge_leduc(torch.from_numpy(kats['leduc_case']).cuda()).cpu().numpy().astype(np.float64)
    assert np.array_equal(got * 2, kats['leduc_payoffs'])


@pytest.mark.gpu
def test_cuda_uno_encoders(kats):
    import torch
    from rlcard_b200 import judgers
    got = judgers.uno_encode(torch.from_numpy(kats['uno_hand']).cuda(), torch.from_numpy(kats['uno_target']).cuda())
    assert np.array_equal(got.cpu().numpy().reshape(-1, 240), kats['uno_obs'])
