"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's DP stage tables, pinned against the reference's own
functions compiled into oracle/_ref/libref_params.so (tests/test_cpu_params.py) and against tests/golden/params_golden.json.

  seed_positions     getSeedPositions                                definitions.h:323-442 (constants :186-240)
  stage_parameters   getParameterFor{Single,Default,NewDefault,Deep}DP   CPUfunctions.cpp:59-260
                     deep DP round 2: maxHitNum 1000                  DV-DPForBothUnalign.cu:138-139
"""
import math

SINGLE, DEFAULT, NEW_DEFAULT, DEEP1, DEEP2 = 1, 2, 3, 4, 5


def _by_length(n, table, last):
    """table: ((bound, value), ...) read as `if n > bound: value`, else last"""
    for bound, value in table:
        if n > bound:
            return value
    return last


def seed_positions(stage, n):
    if stage in (SINGLE, NEW_DEFAULT):                                               # definitions.h:325-376
        seed_len = _by_length(n, ((300, 70), (80, 38), (60, 32), (40, 26)), 22)
        num = 3 + n // 100 if n > 120 else 3
        head = int(n * 0.15) if n > 300 else 0
        trim = int(n * 0.15) if n > 300 else _by_length(n, ((80, 10), (60, 4), (40, 4)), 0)
        apart = int((n - trim - head) / num)                                          # C division, truncating
        pos = [head + i * apart for i in range(num)]
        if pos[-1] > n - seed_len - trim:
            pos[-1] = n - seed_len - trim
        return seed_len, pos
    if stage in (DEEP1, DEEP2):                                                       # :377-441
        if stage == DEEP1:
            seed_len = _by_length(n, ((150, 45), (80, 26), (60, 24), (40, 22)), 20)
        else:
            seed_len = _by_length(n, ((150, 52), (80, 30), (60, 28), (40, 26)), 24)
        head, tail = (int(n * 0.1), int(n * 0.2)) if n > 150 else (0, 0)
        pos = list(range(n - seed_len - tail, head - 1, -(seed_len // 2)))
        if pos and pos[-1] > head:
            pos.append(head)
        return seed_len, pos
    raise ValueError(stage)


def stage_parameters(stage, n1, n2=0, default_threshold=True, threshold=0, front=0, end=0):
    """-> dict(softClipLeft, softClipRight, tailTrimLen, singleDPSeedNum, singleDPSeedPos[3], reads=[{...}, {...}])"""
    reads = [dict(cutoffThreshold=0, maxHitNum=0, sampleDist=0, seedLength=0) for _ in range(2)]
    out = dict(softClipLeft=front, softClipRight=end, tailTrimLen=0, singleDPSeedNum=0, singleDPSeedPos=[0, 0, 0], reads=reads)
    for e, n in enumerate((n1, n2)[:1 if stage == SINGLE else 2]):
        r = reads[e]
        r["cutoffThreshold"] = int(math.ceil(0.3 * float(n))) if default_threshold else threshold
        if stage == DEFAULT:
            r["maxHitNum"] = 50 if n > 50 else 70
        elif stage == NEW_DEFAULT:
            r["maxHitNum"] = 150 if n > 50 else 200
            r["seedLength"] = _by_length(n, ((75, 26), (50, 24)), 22)
        elif stage in (DEEP1, DEEP2):
            r["maxHitNum"] = 1000 if stage == DEEP2 else (100 if n > 50 else 150)
            r["seedLength"] = _by_length(n, ((150, 45), (80, 26), (60, 24), (40, 22)), 20)
            r["sampleDist"] = int(r["seedLength"] * 0.5)
        else:
            r["maxHitNum"] = _by_length(n, ((300, 4), (80, 10), (60, 20), (40, 30)), 40)
            r["seedLength"] = _by_length(n, ((300, 70), (80, 38), (60, 32), (40, 26)), 22)
    if stage == SINGLE:
        out["singleDPSeedNum"] = 3 + n1 // 100 if n1 > 100 else 3
        trim = _by_length(n1, ((80, 10), (60, 4), (40, 4)), 0)
        p2 = int((n1 - trim) * 0.5 - 1)
        if (p2 & 0xFFFFFFFF) > ((n1 - reads[0]["seedLength"]) & 0xFFFFFFFF):           # int against uint: compared unsigned
            p2 = n1 - reads[0]["seedLength"]
        out["singleDPSeedPos"] = [0, int((0 + p2) / 2) - 1, p2]
    return out
