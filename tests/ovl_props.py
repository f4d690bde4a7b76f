"""Size-independent properties of a wtzmo `.ovl` file (17 tab columns, wtzmo.c:1235-1243), checked against the reads it was computed
from.  Used on outputs too large to compare with a CPU run: every record must be internally consistent, inside its reads, above the
-s / -m thresholds, reported from one side only and once per strand, and -- for a sample of records -- its CIGAR is walked over the actual bases: the walk must consume
exactly [tb, te) of the first read and [qb, qe) of the second read on the strand shown, and reproduce the mat / mis / ins / del columns."""
import re

import numpy as np

_CODE = np.full(256, 255, np.uint8)
for _i, _ch in enumerate("ACGT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i
_CIG = re.compile(rb"(\d+)([MID])")


def load_fasta(path):
    """name -> uint8 codes 0..3 (single- or multi-line FASTA, pure ACGT)"""
    reads, name, parts = {}, None, []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    reads[name] = _CODE[np.frombuffer(b"".join(parts), np.uint8)]
                name, parts = line[1:].split()[0], []
            else:
                parts.append(line.strip())
    if name is not None:
        reads[name] = _CODE[np.frombuffer(b"".join(parts), np.uint8)]
    return reads


def check_ovl(reads, ovl_path, min_score=200, min_id=0.6, walk_every=1, dot_matrix=False):
    """returns (records, records walked, aligned columns); raises AssertionError naming the first offending line"""
    seen, n, walked, cols = set(), 0, 0, 0
    with open(ovl_path, "rb") as f:
        for ln, line in enumerate(f, 1):
            c = line.rstrip(b"\n").split(b"\t")
            assert len(c) == 17, (ln, "17 columns expected", len(c))
            q, qs, c1 = reads[c[0]], c[1], reads[c[5]]
            qlen, tb, te, clen, qb, qe, score, mat, mis, ins, dele = (int(c[i]) for i in (2, 3, 4, 7, 8, 9, 10, 12, 13, 14, 15))
            assert qs == b"+" and c[6] in (b"+", b"-"), (ln, "strand columns")
            assert qlen == len(q) and clen == len(c1), (ln, "read lengths differ from the input")
            assert 0 <= tb < te <= qlen and 0 <= qb < qe <= clen, (ln, "coordinates outside the reads")
            assert c[0] != c[5], (ln, "self overlap")
            # a pair is tried from the side of the read processed first and then closed (wtzmo.c:1005-1010), so it never shows up from both
            # sides; the same side may report it once per strand (seeds are per (candidate, strand), hzm_aln.h:896-914)
            assert (c[0], c[5], c[6]) not in seen and (c[5], c[0], b"+") not in seen and (c[5], c[0], b"-") not in seen, (ln, "pair reported twice")
            seen.add((c[0], c[5], c[6]))
            n += 1
            if dot_matrix:
                assert c[16] == b"0M", (ln, "dot-matrix records carry no alignment")
                cols += max(te - tb, qe - qb)
                continue
            aln = mat + mis + ins + dele
            cols += aln
            assert mat + mis + dele == te - tb and mat + mis + ins == qe - qb, (ln, "counts do not add up to the aligned spans")
            assert score >= min_score, (ln, "score below -s")
            assert np.float32(mat) >= np.float32(aln) * np.float32(min_id), (ln, "identity below -m")
            assert c[11] == (b"%0.3f" % (float(mat) / aln)), (ln, "identity column")
            if (n - 1) % walk_every:
                continue
            cs = c1 if c[6] == b"+" else (3 - c1[::-1])
            x1, x2, m2, s2, i2, d2 = tb, qb, 0, 0, 0, 0
            ops = _CIG.findall(c[16])
            assert b"".join(a + b for a, b in ops) == c[16] and ops, (ln, "malformed CIGAR")
            for num, op in ops:
                k = int(num)
                if op == b"M":
                    same = int(np.count_nonzero(q[x1:x1 + k] == cs[x2:x2 + k]))
                    m2 += same; s2 += k - same; x1 += k; x2 += k
                elif op == b"I":
                    i2 += k; x2 += k
                else:
                    d2 += k; x1 += k
            assert (x1, x2) == (te, qe), (ln, "CIGAR does not span the reported coordinates")
            assert (m2, s2, i2, d2) == (mat, mis, ins, dele), (ln, "CIGAR walk over the bases disagrees with the count columns", (m2, s2, i2, d2), (mat, mis, ins, dele))
            walked += 1
    return n, walked, cols
