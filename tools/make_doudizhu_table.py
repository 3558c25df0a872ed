#!/usr/bin/env python
"""Build rlcard_b200/tables/doudizhu_actions.npz from the reference's action-id contract
(games/doudizhu/jsondata.zip: action_space.txt) and cross-check the rule-based classifier in
rlcard_b200/doudizhu_table.py against card_type.json / type_card.json.  Needs /root/reference
(build container only)."""
import io
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlcard_b200 import doudizhu_table as T   # noqa: E402

z = zipfile.ZipFile('/root/reference/rlcard/games/doudizhu/jsondata.zip')
names = {os.path.basename(n): n for n in z.namelist()}
actions = z.read(names['action_space.txt']).decode().split()
card_type = json.loads(z.read(names['card_type.json']).decode())
assert len(actions) == T.NUM_ACTIONS and actions[T.PASS_ID] == 'pass'
packed = np.array([T.pack_counts(T.str_to_counts(a)) for a in actions], np.uint64)
assert len(set(packed.tolist())) == T.NUM_ACTIONS
# canonical spelling: every action string is its rank-sorted form
assert all(T.counts_to_str(T.str_to_counts(a)) == a for a in actions)
np.savez_compressed(os.path.join(ROOT, 'rlcard_b200', 'tables', 'doudizhu_actions.npz'), counts=packed)
# classifier vs the reference tables: same type, and weights ordered identically inside each type
bad = 0
by_type = {}
for a in actions[:-1]:
    ref = card_type[a]
    assert len(ref) == 1, a
    t, w = T.classify(T.str_to_counts(a))
    if t != ref[0][0]:
        bad += 1
        if bad < 10:
            print('type mismatch', a, t, ref)
    by_type.setdefault(t, []).append((int(ref[0][1]), w))
for t, pairs in by_type.items():
    off = {rw - w for rw, w in pairs}
    assert len(off) == 1, (t, off)      # reference weight = ours + constant
print('actions', len(actions), 'type mismatches', bad, 'types', len(by_type))
assert bad == 0
