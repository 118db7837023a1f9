"""Generates tests/golden/mapq_golden.json from the reference's own MAPQ functions (oracle/_ref/libref_mapq.so, built by
oracle/build_ref.sh from BGS-IO.cpp:33-45,2280-2580).  Run in the container that has /root/reference."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import test_cpu_mapq as t  # noqa: E402

ref = t.load_ref()
assert ref is not None, "run oracle/build_ref.sh first"
cs = t.cases(np.random.default_rng(5), n=150)
out = {name: [[list(a), t.reference(ref, name, a)] for a in argl[:400]] for name, argl in cs.items()}
json.dump(dict(source="oracle/_ref/libref_mapq.so (reference code, see oracle/build_ref.sh)", cases=out),
          open(os.path.join(HERE, "mapq_golden.json"), "w"))
print(sum(len(v) for v in out.values()), "rows")
