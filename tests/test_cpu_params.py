"""CPU tier: the DP stage tables.  s3_seed_layout and s3_dp_stage_parameters are host integer tables,
so the product entries run here: against oracle/params_oracle.py, against the reference's own getSeedPositions /
getParameterFor*DP compiled into oracle/_ref/libref_params.so, and against tests/golden/params_golden.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import helpers
from soap3dp_b200 import api

sys.path.insert(0, os.path.join(helpers.ROOT, "oracle"))
import params_oracle as orc  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "params_golden.json")
SEED_STAGES = (1, 3, 4, 5)
LENGTHS = list(range(30, 400)) + [500, 750, 1000]          # DP is off under 30 bases (MIN_READ_LEN_FOR_DP, definitions.h:170)


def product_params(stage, n1, n2, *ini):
    p = api.getParameterForDP(stage, n1, n2, *ini)
    return dict(softClipLeft=p.softClipLeft, softClipRight=p.softClipRight, tailTrimLen=p.tailTrimLen,
                singleDPSeedNum=p.singleDPSeedNum, singleDPSeedPos=list(p.singleDPSeedPos)[:3],
                reads=[dict(cutoffThreshold=r.cutoffThreshold, maxHitNum=r.maxHitNum, sampleDist=r.sampleDist,
                            seedLength=r.seedLength) for r in p.paramRead])


def load_ref():
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_params.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    I = C.POINTER(C.c_int)
    lib.ref_seed_positions.argtypes = [C.c_int, C.c_int, I, I, I]
    lib.ref_stage_parameters.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, I]
    return lib


def ref_seed_positions(lib, stage, n):
    sl, num = C.c_int(0), C.c_int(0)
    pos = (C.c_int * 512)()
    lib.ref_seed_positions(stage, n, C.byref(sl), pos, C.byref(num))
    return sl.value, [pos[i] for i in range(num.value)]


def ref_stage_parameters(lib, stage, n1, n2, default, thr, front, end):
    out = (C.c_int * 22)()
    assert lib.ref_stage_parameters(stage, n1, n2, 1 if default else 0, thr, front, end, out) == 22
    o = list(out)
    reads = [dict(cutoffThreshold=o[14 + 4 * e], maxHitNum=o[15 + 4 * e], sampleDist=o[16 + 4 * e], seedLength=o[17 + 4 * e]) for e in range(2)]
    return dict(softClipLeft=o[0], softClipRight=o[1], tailTrimLen=o[2], singleDPSeedNum=o[3], singleDPSeedPos=o[4:7], reads=reads)


def test_seed_layout_matches_the_restatement():
    for stage in SEED_STAGES:
        for n in LENGTHS:
            assert api.getSeedPositions(stage, n) == orc.seed_positions(stage, n), (stage, n)
    # 100 bp, the bench's read length: three 38-base seeds for single-end DP, 26-base seeds 13 apart for deep DP
    assert api.getSeedPositions(1, 100) == (38, [0, 30, 52])
    sl, pos = api.getSeedPositions(4, 100)
    assert sl == 26 and pos[0] == 74 and pos[-1] == 0 and all(a - b == 13 for a, b in zip(pos[:-2], pos[1:-1]))


def test_seed_layout_edges():
    assert api.getSeedPositions(4, 12) == (20, [])                 # no seed fits (the reference reads seedPositions[-1])
    with pytest.raises(api.S3Error, match="stage 2 has no seeds"):
        api.getSeedPositions(2, 100)
    with pytest.raises(api.S3Error, match="more than 2 seeds"):
        api.getSeedPositions(4, 100, capacity=2)
    with pytest.raises(api.S3Error):
        api.getSeedPositions(1, 0)


def test_stage_parameters_match_the_restatement():
    rng = np.random.default_rng(4)
    for stage in (1, 2, 3, 4, 5):
        for n1 in LENGTHS[::3] + [1, 5, 21, 22, 29]:
            n2 = int(rng.integers(1, 400))
            for ini in ((True, 0, 0, 0), (False, 37, 3, 8)):
                assert product_params(stage, n1, n2, *ini) == orc.stage_parameters(stage, n1, n2, *ini), (stage, n1, n2, ini)
    p = api.getParameterForDP(2, 100, 100)
    assert p.paramRead[0].cutoffThreshold == 30 and p.paramRead[1].maxHitNum == 50
    with pytest.raises(api.S3Error):
        api.getParameterForDP(6, 100)


@pytest.mark.skipif(load_ref() is None, reason="oracle/_ref/libref_params.so not built")
def test_restatement_and_product_match_the_reference_tables():
    ref = load_ref()
    for stage in SEED_STAGES:
        for n in LENGTHS:
            want = ref_seed_positions(ref, stage, n)
            assert want == orc.seed_positions(stage, n) == api.getSeedPositions(stage, n), (stage, n)
    rng = np.random.default_rng(8)
    for stage in (1, 2, 3, 4, 5):
        for n1 in list(range(1, 420)) + [999, 1000, 1001]:
            n2 = int(rng.integers(1, 420))
            for ini in ((True, 0, 0, 0), (False, 41, 2, 9)):
                want = ref_stage_parameters(ref, stage, n1, n2, *ini)
                assert want == orc.stage_parameters(stage, n1, n2, *ini) == product_params(stage, n1, n2, *ini), (stage, n1, n2, ini)


def test_restatement_and_product_match_the_golden_fixture():
    g = json.load(open(GOLDEN))
    assert len(g["seed_positions"]) > 100 and len(g["stage_parameters"]) > 100
    for stage, n, sl, pos in g["seed_positions"]:
        assert (sl, pos) == orc.seed_positions(stage, n) == api.getSeedPositions(stage, n)
    for stage, n1, n2, ini, want in g["stage_parameters"]:
        assert want == orc.stage_parameters(stage, n1, n2, *ini) == product_params(stage, n1, n2, *ini)
