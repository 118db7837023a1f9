"""Generates tests/golden/params_golden.json from the reference's own getSeedPositions / getParameterFor*DP
(oracle/_ref/libref_params.so, built by oracle/build_ref.sh against the reference's unmodified headers).
Run in the container that has /root/reference:  python tests/golden/make_params_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import test_cpu_params as t  # noqa: E402

ref = t.load_ref()
assert ref is not None, "run oracle/build_ref.sh first"
lengths = [30, 35, 36, 40, 41, 50, 51, 60, 61, 75, 76, 80, 81, 100, 101, 120, 121, 125, 150, 151, 200, 250, 300, 301, 400, 1000]
seeds = [[s, n, *t.ref_seed_positions(ref, s, n)] for s in t.SEED_STAGES for n in lengths]
params = [[s, n, m, list(ini), t.ref_stage_parameters(ref, s, n, m, *ini)]
          for s in (1, 2, 3, 4, 5) for n, m in zip(lengths, reversed(lengths)) for ini in ((True, 0, 0, 0), (False, 25, 3, 8))]
json.dump(dict(source="oracle/_ref/libref_params.so (reference code, see oracle/build_ref.sh)", seed_positions=seeds,
               stage_parameters=params), open(os.path.join(HERE, "params_golden.json"), "w"))
print(len(seeds), len(params))
