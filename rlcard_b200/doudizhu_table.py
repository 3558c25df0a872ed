"""DouDizhu action table (host side).

The 27 472 action ids are an interface constant of the reference (games/doudizhu/jsondata.zip ->
action_space.txt, loaded by games/doudizhu/utils.py:21-27); their order is an artefact of how that
file was produced and cannot be derived from the rules, so it is committed as packed rank counts in
``tables/doudizhu_actions.npz`` (made by tools/make_doudizhu_table.py).  Everything else is computed
here from the rules: the (type, weight) of every action (card_type.json / type_card.json in the
reference, utils.py:29-38) and the 54-d action features (envs/doudizhu.py:153-167).
"""
import os

import numpy as np

RANKS = '3456789TJQKA2BR'                       # games/doudizhu/utils.py:41-43
NUM_ACTIONS = 27472
PASS_ID = 27471

# type ids: contiguous id ranges in action_space.txt follow this order
TYPE_NAMES = (['solo', 'pair', 'trio', 'trio_solo', 'trio_pair'] +
              ['solo_chain_%d' % k for k in range(5, 13)] + ['pair_chain_%d' % k for k in range(3, 11)] +
              ['trio_chain_%d' % k for k in range(2, 7)] + ['trio_solo_chain_%d' % k for k in range(2, 6)] +
              ['trio_pair_chain_%d' % k for k in range(2, 5)] + ['four_two_solo', 'four_two_pair', 'bomb', 'rocket', 'pass'])
TYPE_ID = {n: i for i, n in enumerate(TYPE_NAMES)}
T_BOMB, T_ROCKET, T_PASS = TYPE_ID['bomb'], TYPE_ID['rocket'], TYPE_ID['pass']

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tables', 'doudizhu_actions.npz')


def str_to_counts(s):
    c = [0] * 15
    if s != 'pass':
        for ch in s:
            c[RANKS.index(ch)] += 1
    return c


def counts_to_str(c):
    return ''.join(RANKS[i] * c[i] for i in range(15)) or 'pass'


def pack_counts(c):
    return sum(int(v) << (4 * i) for i, v in enumerate(c))


def unpack_counts(x):
    return [(int(x) >> (4 * i)) & 15 for i in range(15)]


def _solo_kickers_ok(kick, start, length):
    """games/doudizhu/judger.py:48-89: no bomb among the kickers, no third copy of a rank adjacent to the
    chain (except '2'), not both jokers; kickers never use a chain rank."""
    if kick[13] and kick[14]:
        return False
    for r in range(15):
        if kick[r] == 0:
            continue
        if start <= r < start + length or kick[r] > 3:
            return False
        if kick[r] == 3 and (r == start - 1 or r == start + length) and r != 12:
            return False
    return True


def classify(c):
    """rank counts -> (type name, weight).  weight orders actions of one type like the reference's
    card_type.json (only the order is observable, utils.py:250-260)."""
    n = sum(c)
    if n == 0:
        return 'pass', 0
    nz = [r for r in range(15) if c[r]]
    lo, hi = nz[0], nz[-1]
    consecutive = (hi - lo + 1 == len(nz)) and hi <= 11
    same = len(set(c[r] for r in nz)) == 1
    if n == 2 and c[13] == 1 and c[14] == 1:
        return 'rocket', 0
    if len(nz) == 1:
        return {1: 'solo', 2: 'pair', 3: 'trio', 4: 'bomb'}[n], lo
    if same and consecutive:
        k, per = len(nz), c[lo]
        if per == 1 and k >= 5:
            return 'solo_chain_%d' % k, lo
        if per == 2 and k >= 3:
            return 'pair_chain_%d' % k, lo
        if per == 3 and k >= 2:
            return 'trio_chain_%d' % k, lo
    found = []
    quads = [r for r in nz if c[r] == 4]
    if n == 6:
        for q in quads:
            kick = list(c); kick[q] -= 4
            if _solo_kickers_ok(kick, q, 1):
                found.append(('four_two_solo', q))
    if n == 8:
        for q in quads:
            kick = list(c); kick[q] -= 4
            if all(v in (0, 2) for v in kick) and sum(1 for v in kick if v == 2) == 2:
                found.append(('four_two_pair', q))
    for k in range(1, 6):
        for kind, per_kick in (('solo', 1), ('pair', 2)):
            if n != (3 + per_kick) * k:
                continue
            for start in range(0, 13 - k + 1 if k == 1 else 12 - k + 1):
                if any(c[r] < 3 for r in range(start, start + k)):
                    continue
                kick = list(c)
                for r in range(start, start + k):
                    kick[r] -= 3
                if kind == 'solo':
                    ok = _solo_kickers_ok(kick, start, k)
                else:
                    ok = all(v in (0, 2) for v in kick) and sum(1 for v in kick if v == 2) == k and not kick[13] and not kick[14]
                if ok:
                    name = ('trio_%s' % kind) if k == 1 else 'trio_%s_chain_%d' % (kind, k)
                    found.append((name, start))
    if len(found) != 1:
        raise ValueError('cannot classify %s: %r' % (counts_to_str(c), found))
    return found[0]


_cache = {}


def load():
    """-> dict(counts u64 [27472] nibble-packed rank counts, type u8, weight u8, length u8, features int8 [27472, 54])"""
    if 'tab' in _cache:
        return _cache['tab']
    if not os.path.exists(_TABLE):
        raise FileNotFoundError('%s missing (run tools/make_doudizhu_table.py where the reference is available)' % _TABLE)
    packed = np.load(_TABLE)['counts'].astype(np.uint64)
    assert packed.shape == (NUM_ACTIONS,)
    types = np.zeros(NUM_ACTIONS, np.uint8); weights = np.zeros(NUM_ACTIONS, np.uint8); length = np.zeros(NUM_ACTIONS, np.uint8)
    feats = np.zeros((NUM_ACTIONS, 54), np.int8)
    for i, x in enumerate(packed):
        c = unpack_counts(x)
        t, w = classify(c)
        types[i], weights[i], length[i] = TYPE_ID[t], w, sum(c)
        feats[i] = cards_to_array(c)
    assert types[PASS_ID] == T_PASS
    _cache['tab'] = dict(counts=packed, type=types, weight=weights, length=length, features=feats)
    return _cache['tab']


def cards_to_array(c):
    """envs/doudizhu.py:153-167 _cards2array on rank counts: 4x13 thermometer, column-major, + 2 joker bits."""
    out = np.zeros(54, np.int8)
    for r in range(13):
        out[4 * r:4 * r + c[r]] = 1
    out[52] = 1 if c[13] else 0
    out[53] = 1 if c[14] else 0
    return out


def action_strings():
    return [counts_to_str(unpack_counts(x)) for x in load()['counts']]


# what every action of a type contains, whatever its kickers: `length` consecutive ranks (within 3..A when length > 1)
# holding at least `mult` cards each; the rocket is both jokers.  (mult, length) per type id; pass: never a candidate.
def type_requirements():
    req = []
    for name in TYPE_NAMES:
        if name in ('solo', 'pair', 'trio'):
            req.append(({'solo': 1, 'pair': 2, 'trio': 3}[name], 1))
        elif name in ('trio_solo', 'trio_pair'):
            req.append((3, 1))
        elif name.startswith('solo_chain_'):
            req.append((1, int(name.rsplit('_', 1)[1])))
        elif name.startswith('pair_chain_'):
            req.append((2, int(name.rsplit('_', 1)[1])))
        elif name.startswith(('trio_chain_', 'trio_solo_chain_', 'trio_pair_chain_')):
            req.append((3, int(name.rsplit('_', 1)[1])))
        elif name in ('four_two_solo', 'four_two_pair', 'bomb'):
            req.append((4, 1))
        elif name == 'rocket':
            req.append((0, 2))
        else:
            req.append((0, 0))
    return req


def type_feasible(counts, t):
    """The device prefilter (game_doudizhu.cuh type_feasibility) on a list of 15 rank counts."""
    mult, length = type_requirements()[t]
    if TYPE_NAMES[t] == 'rocket':
        return counts[13] >= 1 and counts[14] >= 1
    if TYPE_NAMES[t] == 'pass':
        return False
    ok = [c >= mult for c in counts]
    if length == 1:
        return any(ok)
    run = 0
    for r in range(12):                                  # chains live on ranks 3..A
        run = run + 1 if ok[r] else 0
        if run >= length:
            return True
    return False


def build_blob():
    """Device table blob for rlc_upload_tables(RLC_DOUDIZHU, ...) (layout: csrc/tu_doudizhu.cu DdzBlobHeader):
    rows u64[27472] | need u64[2 * (864 + 32)] | type u8[27472] | weight u8[27472] |
    tw_start u32[38][17] (first id of type t with weight >= w; [t][16] = end of type t).
    need holds one pair per mask word j (32 ids): [2j] = nibble-wise min of the 32 rows, [2j+1] = bit t set when some id
    of the word has type t; pairs 864.. describe the batches of 32 mask words (1024 ids) the same way."""
    import struct
    tab = load()
    rows = tab['counts'].astype(np.uint64)
    n_words = (NUM_ACTIONS + 31) // 32
    nib = np.stack([(rows >> np.uint64(4 * r)) & np.uint64(15) for r in range(15)], axis=1).astype(np.uint8)   # [A, 15]
    pad = np.full((n_words * 32 - NUM_ACTIONS, 15), 15, np.uint8)
    mins = np.concatenate([nib, pad]).reshape(n_words, 32, 15).min(axis=1)
    need = np.zeros(2 * (864 + 32), np.uint64)
    bmins = np.concatenate([mins, np.full((864 - n_words, 15), 15, np.uint8)]).reshape(27, 32, 15).min(axis=1)
    for r in range(15):
        need[0:2 * n_words:2] |= mins[:, r].astype(np.uint64) << np.uint64(4 * r)
        need[2 * 864:2 * (864 + 27):2] |= bmins[:, r].astype(np.uint64) << np.uint64(4 * r)
    tbit = np.uint64(1) << tab['type'].astype(np.uint64)
    for i in range(NUM_ACTIONS):
        need[2 * (i >> 5) + 1] |= tbit[i]
        need[2 * (864 + (i >> 10)) + 1] |= tbit[i]
    ntypes = len(TYPE_NAMES)
    tw = np.zeros((ntypes, 17), np.uint32)
    types, weights = tab['type'].astype(np.int64), tab['weight'].astype(np.int64)
    for t in range(ntypes):
        ids = np.nonzero(types == t)[0]
        lo, hi = int(ids.min()), int(ids.max()) + 1
        assert np.array_equal(ids, np.arange(lo, hi)), 'type ids must be contiguous'
        w = weights[lo:hi]
        assert np.all(np.diff(w) >= 0), 'weights must be non-decreasing inside a type'
        for k in range(16):
            tw[t, k] = lo + int(np.searchsorted(w, k, side='left'))
        tw[t, 16] = hi
    parts = [rows.tobytes(), need.tobytes(), tab['type'].tobytes(), tab['weight'].tobytes(), tw.tobytes()]
    hdr_size = 64
    offs, cur = [], hdr_size
    for b in parts:
        cur = (cur + 15) & ~15
        offs.append(cur)
        cur += len(b)
    total = (cur + 15) & ~15
    hdr = struct.pack('<4sIII6Q', b'DDZ2', NUM_ACTIONS, 864 + 32, ntypes, *offs, total)
    blob = bytearray(total)
    blob[:len(hdr)] = hdr
    for o, b in zip(offs, parts):
        blob[o:o + len(b)] = b
    return bytes(blob)
