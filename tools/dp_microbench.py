#!/usr/bin/env python
"""GPU micro-benchmark of the banded extension executors through the C ABI (zmo_dp_extend):
row latency of ONE long job (the tail of a DP phase) and throughput of a machine-filling batch, per band class."""
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartdenovo_b200 import Zmo  # noqa: E402
from smartdenovo_b200.api import DP_PROBLEM  # noqa: E402


def mutate(rng, s, sub=0.03, ins=0.07, dele=0.05):
    out = []
    for b in s:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + sub:
            out.append((b + 1 + rng.integers(0, 3)) % 4)
        else:
            out.append(b)
        if rng.random() < ins:
            out.append(rng.integers(0, 4))
    return np.array(out, np.uint8)


def main():
    rng = np.random.default_rng(7)
    L = 12000
    base = rng.integers(0, 4, L).astype(np.uint8)
    reads = [base, mutate(rng, base)]
    z = Zmo()
    z.upload_seqs(reads)
    out = []
    for name, ew, rows in [("cls3_1601", 800, 8000), ("cls2_801", 400, 8000), ("cls1_401", 200, 8000), ("cls0_201", 100, 8000)]:
        for n in (1, 148 * 3, 148 * 12):
            probs = np.zeros(n, DP_PROBLEM)
            qlen = min(rows, len(reads[1]))
            tlen = min(rows, len(reads[0]))
            for i in range(n):
                probs[i] = (1, 0, 0, 1, 0, qlen, 0, 1, 0, tlen, 1000, -ew)
            z.dp_extend(1, probs)      # warm-up
            c0, s0 = z.counters()["cells_ext"], z.stage_ms()["end_extend"]
            t0 = time.perf_counter()
            res, _ = z.dp_extend(1, probs)
            wall = time.perf_counter() - t0
            c1, s1 = z.counters()["cells_ext"], z.stage_ms()["end_extend"]
            ms = s1 - s0
            cells = c1 - c0
            rows_swept = cells / n / (2 * ew + 1)
            out.append(dict(cls=name, jobs=n, kernel_ms=ms, wall_ms=1e3 * wall, cells=cells, gcells_per_s=cells / ms / 1e6,
                            approx_rows_per_job=rows_swept, us_per_row=1e3 * ms / max(1.0, rows_swept) if n == 1 else None,
                            qe=int(res[0]["qe"]), te=int(res[0]["te"]), score=int(res[0]["score"])))
            print(json.dumps(out[-1]), flush=True)
    z.close()


if __name__ == "__main__":
    main()
