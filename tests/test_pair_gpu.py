"""GPU tier: s3_pair_occurrences (paired-end pairing of two occurrence lists, batched over read pairs) against the pairing
oracle, which the CPU tier pins against the reference's PEMappingOccurrences / PEStatsPEPairList.  The per-read-pair
walk both kernels run is also checked on the CPU tier from the same source (tests/test_cpu_pair_walk.py).  The same
comparisons were first run on a B200 through tools/pair_gpu_check.py (profiles/r03b_pair_retain_check.txt)."""
import numpy as np
import pytest

import helpers
from soap3dp_b200 import api, fmindex, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def gi():
    idx = fmindex.build_index(synth.random_genome(50_000, seed=3))
    h = api.GPUINDEXUpload(idx, device=0)
    yield h
    try:
        api.GPUINDEXFree(h)
    except Exception:          # noqa: BLE001 -- a fault in the entry under test must not turn into a teardown error
        pass


@pytest.mark.parametrize("legs", [(1, 2), (2, 1), (1, 1)])
@pytest.mark.parametrize("report_one", [False, True])
def test_pairs_match_the_oracle(gi, legs, report_one):
    rng = np.random.default_rng(23 + legs[0] * 2 + legs[1] + int(report_one))
    for near_edges, npairs in ((False, 3000), (True, 500)):
        lists = helpers.make_occurrence_lists(rng, npairs, max_occ=16, near_edges=near_edges)
        pl = rng.integers(60, 151, npairs).astype(np.uint32)
        got = api.pair_occurrences(gi, *lists, pl, 200, 500, *legs, report_one)
        want = helpers.oracle_pair_occurrences(lists, pl, 200, 500, *legs, report_one)
        assert helpers.same_pairing(got, want)
        assert len(want["pos1"]) > 0


def test_pairs_of_empty_lists_and_batches(gi):
    z32, z8 = np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    got = api.pair_occurrences(gi, z32, z8, z8, np.zeros(1, np.uint64), z32, z8, z8, np.zeros(1, np.uint64), z32, 200, 500)
    assert got["offsets"].tolist() == [0] and len(got["pos1"]) == 0
    z64 = np.zeros(4, np.uint64)
    got = api.pair_occurrences(gi, z32, z8, z8, z64, z32, z8, z8, z64, np.full(3, 100, np.uint32), 200, 500)
    assert got["offsets"].tolist() == [0, 0, 0, 0] and (got["optimal"] == 0xFFFFFFFF).all()
    rng = np.random.default_rng(2)
    lists = list(helpers.make_occurrence_lists(rng, 200))
    lists[4:] = [z32, z8, z8, np.zeros(201, np.uint64)]                      # the second reads have no hits at all
    got = api.pair_occurrences(gi, *lists, np.full(200, 100, np.uint32), 200, 500)
    assert int(got["offsets"][-1]) == 0


@pytest.mark.parametrize("mode,cap", [(0, 0), (1, 1), (1, 3), (1, 40), (2, 0)])
def test_best_hit_filters_match_the_oracle(gi, mode, cap):
    """s3_retain_best against oracle/retain_oracle.c (pinned against the reference's retainAllBest family on the CPU tier)"""
    rng = np.random.default_rng(50 + 7 * mode + cap)
    for nreads, max_sa, max_occ in ((20000, 7, 7), (3000, 9, 2), (3000, 2, 9), (3, 1, 1)):
        lists = helpers.make_hit_lists(rng, nreads, max_sa=max_sa, max_occ=max_occ)
        got = api.retain_best(gi, mode, *lists, cap)
        want = helpers.oracle_retain_best(lists, mode, cap)
        assert helpers.same_retained(got, want), (mode, cap, nreads)


def test_best_hit_filter_rejects_bad_arguments(gi):
    lists = helpers.make_hit_lists(np.random.default_rng(1), 10)
    with pytest.raises(api.S3Error, match="mode"):
        api.retain_best(gi, 3, *lists)
    with pytest.raises(api.S3Error, match="maxNum"):
        api.retain_best(gi, 1, *lists, 0)
