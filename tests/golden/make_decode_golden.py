"""Generates tests/golden/decode_golden.json from the reference's own decoder (oracle/_ref/libref_decode.so, built by
oracle/build_ref.sh from DV-DPfunctions.h:514-597, DV-DPfunctions.cu:1699-1733, PE.cpp:83-110,420-485).  Run in the
container that has /root/reference:  python tests/golden/make_decode_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import helpers  # noqa: E402
from soap3dp_b200 import synth  # noqa: E402

ref = helpers.load_ref_decode()
assert ref is not None, "run oracle/build_ref.sh first"
rng = np.random.default_rng(77)
cases = []
G = synth.random_genome(120_000, seed=5)
for kind, scores4 in (("single", (1, -2, -3, -1)), ("rescue", (2, -3, -5, -2)), ("synthetic", (1, -1, -2, -1))):
    if kind == "synthetic":
        n, pat_len, cutoff = 150, 96, 5
        pat = helpers.synthetic_patterns(rng, n, pat_len, max_ops=24)
        sc = rng.integers(-20, 120, n).astype(np.int32)
        hit = rng.integers(0, 300, n).astype(np.uint32)
        cnt = np.ones(n, np.uint32)
        ln = rng.integers(20, 150, n).astype(np.uint32)
    else:
        b = helpers.make_dp_batch(G, 120, 100, kind, seed=4, indel_rate=0.012)
        sc, hit, cnt, pat, _ = helpers.oracle_dp(helpers.load_oracle_dp(), b, scores4)
        n, pat_len, cutoff, ln = b.n, b.pat_len, 30, b.read_len
        sc, hit, cnt, pat = sc[:n], hit[:n], cnt[:n], pat[:n * pat_len]
    res = helpers.ref_decode(ref, pat, pat_len, sc, hit, ln, np.zeros(n, np.uint32), cnt, cutoff, scores4)
    cases.append(dict(kind=kind, scores4=list(scores4), cutoff=cutoff, pattern_length=int(pat_len),
                      patterns_hex=[bytes(pat[t * pat_len:(t + 1) * pat_len]).rstrip(b'\0').hex() for t in range(n)],   # zero padding dropped
                       scores=[int(x) for x in sc], read_lengths=[int(x) for x in ln],
                      results=[[r[0], r[2], r[3], r[4]] for r in res]))
json.dump(dict(source="oracle/_ref/libref_decode.so (reference code, see oracle/build_ref.sh)", cases=cases),
          open(os.path.join(HERE, "decode_golden.json"), "w"))
print("wrote", sum(len(c["results"]) for c in cases), "results")
