import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
from conftest import mutate
from smartdenovo_b200 import Zmo
from smartdenovo_b200.api import PAIR, TASK
rng = np.random.default_rng(11)
genome = rng.integers(0, 4, 50000).astype(np.uint8)
reads = []
for _ in range(100):
    s = int(rng.integers(0, 50000 - 9000))
    r = mutate(rng, genome[s:s + int(rng.integers(6000, 9000))])
    reads.append((3 - r[::-1]).astype(np.uint8) if rng.random() < 0.5 else r)
reads.sort(key=len, reverse=True)
z = Zmo(); z.upload_seqs(reads); z.index_build()
off, ev = z.candidates(np.arange(len(reads), dtype=np.uint32))
pairs = []
for q in range(len(reads)):
    for e in ev[int(off[q]):int(off[q + 1])]:
        c = int(e["tkey"]) >> 1
        if c != q and (q, c) not in pairs:
            pairs.append((q, c))
pairs = np.array(pairs, PAIR)
def cmp(tag, a, b):
    if a.shape != b.shape:
        print(tag, "shape", a.shape, b.shape); return
    for f in a.dtype.names:
        d = np.nonzero(np.atleast_1d((a[f] != b[f]).reshape(len(a), -1).any(axis=1)))[0]
        if len(d): print(tag, f, "differs at", len(d), "rows, first", d[:5], a[f][d[:3]], b[f][d[:3]])
s0, w0 = z.pair_windows(pairs)
s0b, w0b = z.pair_windows(pairs)
cmp("root-root seeds", s0, s0b); cmp("root-root wins", w0, w0b)
c = z.clone()
s1, w1 = c.pair_windows(pairs)
cmp("root-clone seeds", s0, s1); cmp("root-clone wins", w0, w1)
s2, w2 = c.pair_windows(pairs)
cmp("clone-clone seeds", s1, s2); cmp("clone-clone wins", w1, w2)
print("done", len(pairs), len(w0))
