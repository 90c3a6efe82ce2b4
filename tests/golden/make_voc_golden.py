#!/usr/bin/env python
"""Fixture from the reference's SHIPPED vocabulary (Vocabulary/orbvoc.dbow3: K = 10, L = 6, 1 082 073 nodes, TF_IDF / L1, quicklz
compressed; 49 MB, cannot travel): the part of the real tree that a fixed set of real ORB descriptors walks through.

  1. the reference (oracle/_ref: DBoW3 compiled unmodified) loads orbvoc.dbow3 with its own loader and writes it back uncompressed
     (Vocabulary::save); oracle.ref.read_dbow3_binary turns that stream into flat arrays;
  2. 260 ORB descriptors of a synthetic scene (oracle extraction) are descended through the real tree; the sub-vocabulary keeps the
     root and ALL children of every node on one of those paths, with their real descriptors and weights (nodes renumbered in
     breadth-first order, words in their original order; a pruned inner node becomes a word of weight 0.5 so that the tree stays a
     valid DBoW3 vocabulary for descriptors that wander off the recorded paths);
  3. the expected BowVector / FeatureVector are what the REFERENCE computes (Object::ComputeBow -> Vocabulary::transform, levelsup 4)
     after loading the sub-vocabulary through DBoW3's loader — for the 260 descriptors plus 60 random ones. The script also checks
     that the 260 real descriptors get the same word weights from the sub-vocabulary as from the full one.

Run in the build container (needs /root/reference):  python tests/golden/make_voc_golden.py
"""
import os, sys, tempfile
from collections import deque
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O, ref as R
from mcvslam_b200 import synth

R.build(); O.build()
tmp = tempfile.mkdtemp()
assert R.voc_load("/root/reference/Vocabulary/orbvoc.dbow3") == 971814
assert R.voc_save_uncompressed(os.path.join(tmp, "full.bin")) == 0
V = R.read_dbow3_binary(os.path.join(tmp, "full.bin"))
co, ci, nd, wi, ww = V["child_off"], V["child_ids"], V["node_desc"], V["word_id"], V["weight"]

n, k, d = O.Orb(2000, 1.2, 8, 28, 15).extract(synth.scene(77))
real = d[:: max(1, n // 260)][:260].copy()
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)
keep = {0}
for x in real:                                              # Vocabulary::transform's descent: first minimum wins
    node = 0
    while co[node + 1] > co[node]:
        kids = ci[co[node]:co[node + 1]]
        keep.update(int(c) for c in kids)
        node = int(kids[np.argmin(POP[nd[kids] ^ x].sum(1))])
# breadth-first renumbering, siblings in stored order
new_id = {0: 0}; order = [0]; dq = deque([0])
while dq:
    p = dq.popleft()
    for c in ci[co[p]:co[p + 1]]:
        c = int(c)
        if c in keep:
            new_id[c] = len(order); order.append(c); dq.append(c)
m = len(order)
children = [[] for _ in range(m)]
par = np.zeros(len(wi), np.int64); par[ci] = np.repeat(np.arange(len(wi)), np.diff(co))
for old in order[1:]:
    children[new_id[int(par[old])]].append(new_id[old])
s_co = np.zeros(m + 1, np.int32); s_ci = []
for i in range(m):
    s_co[i + 1] = s_co[i] + len(children[i]); s_ci += children[i]
s_nd = nd[order].copy()
leaf = np.diff(s_co) == 0
s_w = np.where(leaf, np.where(wi[order] >= 0, ww[order], 0.5), 0.0)
# words: real words first in their original order, then the pruned inner nodes
is_word = leaf & (wi[order] >= 0)
rank = np.full(m, -1, np.int32)
real_words = np.nonzero(is_word)[0]
rank[real_words[np.argsort(wi[np.array(order)[real_words]], kind="stable")]] = np.arange(len(real_words))
pruned = np.nonzero(leaf & ~is_word)[0]
rank[pruned] = len(real_words) + np.arange(len(pruned))
sub = dict(child_off=s_co, child_ids=np.array(s_ci, np.uint32), node_desc=s_nd, word_id=rank, weight=s_w.astype(np.float64), L=V["L"], K=V["K"],
           weighting=V["weighting"], norm=V["norm"], scoring=V["scoring"])

desc = np.concatenate([real, synth.descriptors(60, 9)])
kp = np.zeros(len(desc), O.KP_DTYPE)
full = R.Obj(R.Orb(), np.zeros(len(real), O.KP_DTYPE), real, 640, 480).compute_bow()        # with the FULL vocabulary still loaded
assert R.voc_load(R.write_dbow3_binary(sub, os.path.join(tmp, "sub.dbow3"))) == int((rank >= 0).sum())
part = R.Obj(R.Orb(), np.zeros(len(real), O.KP_DTYPE), real, 640, 480).compute_bow()
assert np.array_equal(np.sort(full["bow_vals"]), np.sort(part["bow_vals"])), "sub-vocabulary changes the real descriptors' word weights"
exp = R.Obj(R.Orb(), kp, desc, 640, 480).compute_bow()
o = O.bow_transform(desc, sub, 4)
assert np.array_equal(exp["bow_ids"], o["bow_ids"]) and exp["bow_vals"].tobytes() == o["bow_vals"].tobytes()
out = os.path.join(ROOT, "tests", "golden", "voc_golden.npz")
np.savez_compressed(out, desc=desc, n_real=len(real), **{"voc_" + a: np.asarray(b) for a, b in sub.items()}, **{"exp_" + a: b for a, b in exp.items()})
print("nodes %d (of %d), words %d, %d distinct words hit, %.0f KB" % (m, len(wi), int((rank >= 0).sum()), len(exp["bow_ids"]), os.path.getsize(out) / 1024))
