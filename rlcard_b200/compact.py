"""Host side of the compact trajectory wire format (csrc/tu_compact.cu, rlc_compact_trajectory): expansion of the packed
records into the dense arrays ``VecEnv.rollout_random`` would have delivered.  numpy only -- this runs on the consumer's
side of PCIe."""
import numpy as np


def expand(game, packed):
    """packed: uint32 array [..., words] (torch CPU tensor or numpy) -> dict of dense numpy arrays with the leading
    shape of ``packed``: obs uint8 [..., D] (Scout: float32), mask uint8 [..., A] (DouDizhu: the bit-packed uint32 [..., 860] rows of
    the dense format), action / player int32, done uint8, payoffs float32 [..., P]."""
    p = np.asarray(packed).view(np.uint32)
    lead = p.shape[:-1]
    if game == 'leduc-holdem':
        w = p[..., 0]
        obs = np.zeros((w.size, 36), np.uint8)
        f = w.reshape(-1)
        rows = np.arange(f.size)
        obs[rows, (f & 3).astype(np.int64)] = 1                          # own card
        pub = ((f >> 2) & 3).astype(np.int64)
        dealt = pub > 0
        obs[rows[dealt], 2 + pub[dealt]] = 1                              # row[3 + public rank]
        obs[rows, 6 + ((f >> 4) & 15).astype(np.int64)] = 1               # own chips
        obs[rows, 21 + ((f >> 8) & 15).astype(np.int64)] = 1              # the other seat's chips
        obs = obs.reshape(lead + (36,))
        meta, scale = w, 0.25
    elif game == 'limit-holdem':
        cards = p[..., 0].astype(np.uint64) | (p[..., 1].astype(np.uint64) << np.uint64(32))
        obs = np.zeros(lead + (72,), np.uint8)
        obs[..., :52] = ((cards[..., None] >> np.arange(52, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)
        meta = p[..., 2]
        flat = obs.reshape(-1, 72)
        rows = np.arange(flat.shape[0])
        for r in range(4):
            flat[rows, (52 + 5 * r + ((meta.reshape(-1) >> (3 * r)) & 7)).astype(np.int64)] = 1
        scale = 0.5
    elif game == 'uno':
        return _expand_uno(p, lead)
    elif game == 'doudizhu':
        return _expand_doudizhu(p, lead)
    elif game == 'scout':
        return _expand_scout(p, lead)
    else:
        raise ValueError('no compact wire format for %r' % game)
    nib = (meta >> 12) & 15
    mask = ((nib[..., None] >> np.arange(4, dtype=np.uint32)) & 1).astype(np.uint8)
    q = ((meta >> 20) & 255).astype(np.uint8).view(np.int8).astype(np.float32) * np.float32(scale)
    return {'obs': obs, 'mask': mask, 'action': ((meta >> 16) & 3).astype(np.int32), 'player': ((meta >> 18) & 1).astype(np.int32),
            'done': ((meta >> 19) & 1).astype(np.uint8), 'payoffs': np.stack([q, -q], axis=-1)}


def _bits(words, n):
    """uint32 [m, k] -> uint8 [m, n]: bit j of the concatenated words."""
    j = np.arange(n)
    return ((words[:, j >> 5] >> (j & 31).astype(np.uint32)) & 1).astype(np.uint8)


def _expand_uno(p, lead):
    f = p.reshape(-1, 7)
    m = f.shape[0]
    rows = np.arange(m)
    k = _bits(f[:, 0:2], 60).astype(np.int64) + 2 * _bits(f[:, 2:4], 60).astype(np.int64)      # copies held per (colour, trait): 0, 1, 2
    obs = np.zeros((m, 240), np.uint8)
    pos = np.arange(60)
    obs[rows[:, None], 60 * k + pos[None, :]] = 1                                               # envs/uno.py:24-33 planes 0..2
    meta = f[:, 6]
    obs[rows, 180 + ((meta >> 6) & 63).astype(np.int64)] = 1                                    # plane 3: the target
    win = (meta >> 20) & 3
    p0 = np.where(win == 1, 1.0, np.where(win == 2, -1.0, 0.0)).astype(np.float32)
    return {'obs': obs.reshape(lead + (240,)), 'mask': _bits(f[:, 4:6], 61).reshape(lead + (61,)),
            'action': (meta & 63).astype(np.int32).reshape(lead), 'player': ((meta >> 16) & 1).astype(np.int32).reshape(lead),
            'done': ((meta >> 19) & 1).astype(np.uint8).reshape(lead), 'payoffs': np.stack([p0, -p0], axis=-1).reshape(lead + (2,))}


_ddz = {}


def doudizhu_legal_mask(hand, target):
    """Bit-packed legal set (uint32 [860]) of a DouDizhu decision from the two things it depends on: the 15 rank counts of the
    current hand and of the action to beat (all zero = lead).  games/doudizhu: player.py:60-76 available_actions over
    judger.py:124-331 (lead: every non-pass action contained in the hand) and utils.py:225-262 get_gt_cards (follow: 'pass' +
    contained actions of the target's type with a larger weight + bombs, larger bombs / the rocket against a bomb, only
    'pass' against the rocket) -- the same statement as csrc/game_doudizhu.cuh legal(), over the same action table."""
    from . import doudizhu_table as T
    if not _ddz:
        tab = T.load()
        _ddz.update(rows=tab['counts'], type=tab['type'].astype(np.int64), weight=tab['weight'].astype(np.int64),
                    ids={int(x): i for i, x in enumerate(tab['counts'])})
    M = np.uint64(0x8888888888888888)
    H = np.uint64(T.pack_counts(hand))
    ok = (((H | M) - _ddz['rows']) & M) == M
    ok[T.PASS_ID] = False
    tgt = T.pack_counts(target)
    if tgt != 0:
        t = _ddz['ids'][tgt]
        tt, tw = _ddz['type'][t], _ddz['weight'][t]
        ty, we = _ddz['type'], _ddz['weight']
        if tt == T.T_ROCKET:
            ok[:] = False
        elif tt == T.T_BOMB:
            ok &= ((ty == T.T_BOMB) & (we > tw)) | (ty == T.T_ROCKET)
        else:
            ok &= ((ty == tt) & (we > tw)) | (ty == T.T_BOMB) | (ty == T.T_ROCKET)
        ok[T.PASS_ID] = True
    bits = np.zeros(860 * 32, np.uint8)
    bits[:T.NUM_ACTIONS] = ok
    return np.packbits(bits, bitorder='little').view(np.uint32)


def _expand_doudizhu(p, lead):
    f = p.reshape(-1, 33)
    m = f.shape[0]
    meta = f[:, 32]
    seat = ((meta >> 16) & 3).astype(np.int64)
    counts = np.zeros((m, 16, 15), np.int64)                         # rank counts of the 16 blocks
    for r in range(15):
        counts[:, :, r] = (f[:, (0 if r < 8 else 1):32:2] >> np.uint32(4 * (r & 7))) & 15
    obs = np.zeros((m, 912), np.uint8)
    blocks = obs[:, :864].reshape(m, 16, 54)
    for r in range(13):                                              # envs/doudizhu.py:153-167: 4 x 13 thermometer, then the jokers
        for k in range(4):
            blocks[:, :, 4 * r + k] = counts[:, :, r] > k
    blocks[:, :, 52] = counts[:, :, 13] > 0
    blocks[:, :, 53] = counts[:, :, 14] > 0
    rows = np.arange(m)
    pa, pb = ((meta >> 21) & 31).astype(np.int64), ((meta >> 26) & 31).astype(np.int64)
    lord = seat == 0
    obs[rows, np.where(lord, 756, 864) + pa] = 1                     # cards-left one-hots (landlord: 17 + 17, peasants: 20 + 17)
    obs[rows, np.where(lord, 773, 884) + pb] = 1
    mask = np.zeros((m, 860), np.uint32)
    memo = {}
    for i in range(m):                                               # the legal set depends on (hand, action to beat) only
        key = (int(f[i, 0]), int(f[i, 1]), int(f[i, 4]), int(f[i, 5]))
        if key not in memo:
            memo[key] = doudizhu_legal_mask(counts[i, 0], counts[i, 2])
        mask[i] = memo[key]
    won = ((meta >> 20) & 1).astype(np.float32)
    done = ((meta >> 19) & 1).astype(np.uint8)
    pay = np.stack([won, (1 - won) * done, (1 - won) * done], axis=-1).astype(np.float32)
    return {'obs': obs.reshape(lead + (912,)), 'mask': mask.reshape(lead + (860,)), 'action': (meta & 0x7fff).astype(np.int32).reshape(lead),
            'player': seat.astype(np.int32).reshape(lead), 'done': done.reshape(lead), 'payoffs': pay.reshape(lead + (3,))}


def _expand_scout(p, lead):
    f = p.reshape(-1, 20)
    m = f.shape[0]
    rows = np.arange(m)
    obs = np.zeros((m, 688), np.float32)
    slot = np.arange(16)
    tops = None
    for plane in range(4):                                           # hand tops / bottoms, table tops / bottoms: 16 slots x 10 values
        val = ((f[:, 2 * plane + (slot >> 3)] >> (4 * (slot & 7)).astype(np.uint32)) & 15).astype(np.int64)
        if plane == 2:
            tops = val
        r, c = np.nonzero(val)
        obs[r, 160 * plane + 10 * c + val[r, c] - 1] = 1
        if plane == 0:
            obs[r, 640 + c] = 1
        if plane == 2:
            obs[r, 656 + c] = 1
    meta = f[:, 19]
    owner, consec, tl = ((meta >> 23) & 7).astype(np.int64), (meta >> 21) & 3, ((meta >> 26) & 31).astype(np.int64)
    obs[rows, 672 + owner] = 1                                       # envs/scout.py:171-235 scalars (float32 of the float64 expressions)
    third = np.array([0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0]).astype(np.float32)
    obs[:, 677] = third[consec]
    for k in range(4):
        obs[:, 678 + k] = ((f[:, 15] >> np.uint32(5 * k)) & 31).astype(np.float32) * np.float32(0.0625)
    obs[:, 682] = f[:, 16].astype(np.float32) * np.float32(0.0625)
    obs[:, 683] = tl.astype(np.float32) * np.float32(0.0625)
    obs[:, 685] = ((meta >> 20) & 1).astype(np.float32)
    tenth = (np.arange(11) / 10.0).astype(np.float32)
    first = np.where(tl > 0, tops[:, 0], 0)
    last = np.where(tl > 0, tops[rows, np.maximum(tl - 1, 0)], 0)
    obs[:, 686], obs[:, 687] = tenth[first], tenth[last]
    pay = np.stack([(f[:, 17] & 0xffff).astype(np.uint16).view(np.int16), (f[:, 17] >> 16).astype(np.uint16).view(np.int16),
                    (f[:, 18] & 0xffff).astype(np.uint16).view(np.int16), (f[:, 18] >> 16).astype(np.uint16).view(np.int16)], axis=-1)
    return {'obs': obs.reshape(lead + (688,)), 'mask': _bits(f[:, 8:15], 204).reshape(lead + (204,)),
            'action': (meta & 255).astype(np.int32).reshape(lead), 'player': ((meta >> 16) & 3).astype(np.int32).reshape(lead),
            'done': ((meta >> 19) & 1).astype(np.uint8).reshape(lead), 'payoffs': pay.astype(np.float32).reshape(lead + (4,))}
