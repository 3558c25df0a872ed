"""Host side of the compact trajectory wire format (csrc/tu_compact.cu, rlc_compact_trajectory): expansion of the packed
records into the dense arrays ``VecEnv.rollout_random`` would have delivered.  numpy only -- this runs on the consumer's
side of PCIe."""
import numpy as np


def expand(game, packed):
    """packed: uint32 array [..., words] (torch CPU tensor or numpy) -> dict of dense numpy arrays with the leading
    shape of ``packed``: obs uint8 [..., D], mask uint8 [..., 4], action / player int32, done uint8, payoffs float32 [..., 2]."""
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
    else:
        raise ValueError('no compact wire format for %r' % game)
    nib = (meta >> 12) & 15
    mask = ((nib[..., None] >> np.arange(4, dtype=np.uint32)) & 1).astype(np.uint8)
    q = ((meta >> 20) & 255).astype(np.uint8).view(np.int8).astype(np.float32) * np.float32(scale)
    return {'obs': obs, 'mask': mask, 'action': ((meta >> 16) & 3).astype(np.int32), 'player': ((meta >> 18) & 1).astype(np.int32),
            'done': ((meta >> 19) & 1).astype(np.uint8), 'payoffs': np.stack([q, -q], axis=-1)}
