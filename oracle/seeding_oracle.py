"""TEST INFRASTRUCTURE ONLY: the DP stages that start from seeds, composed on the host from the other oracles (search slots,
seed merges, window selection, DP, CIGAR encoder -- each pinned against the reference's own code, see their headers) and
from restatements of the reference's host steps between them.  Only tests/ and bench.py's checker legs may import it.

Follows:
  single_1_mismatch_alignment2 + hostKernelSingle   alignment.cu:1839-1893, CPUfunctions.cpp:2700-2870   seeding driver
  PairEndSeedingBatch::packSeeds / pack             DV-DPfunctions.cu:2655-2706   seeds of both mates of a pair
  PairEndSeedingEngine::performSeeding              DV-DPfunctions.cu:3025-3160   seeded / too-many / unseeded pairs
  DeepDPWrapper::seeding_ext                        DV-DPForBothUnalign.cu:131-143   round 1, then round 2 for the too-many pairs
  PairEndAlignmentEngine (packLeft, align, packRight, align)   DV-DPfunctions.cu:3374-3472,3574-3800
  SingleEndSeedingEngine / SingleEndAlignmentEngine  DV-DPfunctions.cu:1026-1370,1373-1780  (single-read DP)
The round logic and the result assembly have no reference build to run against (the engines need the whole aligner around
them): parity of those steps is unpinned and DESIGN.md says so; every arithmetic piece underneath is pinned.
"""
import numpy as np


def complete_ranges(launch, n, k, num_cases, allowed=4096):
    """launch(case, k, allowed, wpa) -> answers view [n, wpa]; -> per read the uncapped list of (l, r, strand) in slot order"""
    wpa = 2 * allowed
    out = [[] for _ in range(n)]
    for case in range(num_cases):
        v = launch(case, k, allowed, wpa)
        for q in range(n):
            row = v[q]
            assert int(row[0]) != 0xFFFFFFFE, "oracle slot overflowed; enlarge it"
            if int(row[0]) == 0xFFFFFFFD:
                continue
            for s in range(allowed):
                a0, a1 = int(row[2 * s]), int(row[2 * s + 1])
                if a0 == 0xFFFFFFFF and a1 == 0xFFFFFFFF:
                    break
                out[q].append((a0, a0 + (a1 & 0xFFFFFF), ((a1 >> 27) & 1) + 1))
    return out


def seeding_driver(launch_exact, launch_one, n, max_hit):
    """exact search of every seed, 1-mismatch search of those without an alignment, maxHitNum cap.
    launch_exact(case, k, allowed, wpa) over all n seeds; launch_one(ids)(case, k, allowed, wpa) over the seeds `ids`.
    -> per seed (status, ranges)"""
    exact = complete_ranges(launch_exact, n, 0, 1)
    ids = [s for s in range(n) if not exact[s]]
    one = complete_ranges(launch_one(ids), len(ids), 1, 2) if ids else []
    second = dict(zip(ids, one))
    out = []
    for s in range(n):
        ranges = second[s] if s in second else exact[s]
        occ = sum(r - l + 1 for l, r, _ in ranges)
        mh = int(max_hit[s]) if hasattr(max_hit, "__len__") else int(max_hit)
        if occ == 0:
            out.append((0, []))
        elif occ <= mh:
            out.append((1, ranges))
        else:
            out.append((4, []))
    return out
