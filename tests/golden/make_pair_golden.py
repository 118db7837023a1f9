"""Generates tests/golden/pair_golden.json from the reference's own PEMappingOccurrences + PEStatsPEOutput
(oracle/_ref/libref_pair.so, built by oracle/build_ref.sh from PEAlgnmt.cpp).  Run in the container that has
/root/reference:  python tests/golden/make_pair_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import helpers  # noqa: E402

ref = helpers.load_ref_pair()
assert ref is not None, "run oracle/build_ref.sh first"
rng = np.random.default_rng(99)
cases = []
for legs, report_one, near_edges, bounds in (((1, 2), False, False, (200, 500)), ((2, 1), False, True, (1, 300)),
                                             ((1, 1), True, False, (200, 500)), ((1, 2), True, True, (200, 500))):
    lists = helpers.make_occurrence_lists(rng, 60, max_occ=9, near_edges=near_edges)
    pl = rng.integers(60, 151, 60).astype(np.uint32)
    want = helpers.ref_pair_occurrences(ref, lists, pl, *bounds, *legs, report_one)
    cases.append(dict(legs=list(legs), report_one=report_one, bounds=list(bounds), pattern_lengths=pl.tolist(),
                      lists=[v.tolist() for v in lists],
                      want={k: np.asarray(v).tolist() for k, v in want.items() if k != "stats"},
                      stats_nonzero=[[int(p), int(k), int(want["stats"][p, k])] for p, k in zip(*np.nonzero(want["stats"]))]))
json.dump(dict(source="oracle/_ref/libref_pair.so (reference code, see oracle/build_ref.sh)", cases=cases),
          open(os.path.join(HERE, "pair_golden.json"), "w"))
print(sum(len(c["want"]["pos1"]) for c in cases), "pairs")
