"""Host-side buffer formats of the reference's GPU boundary.

Only numpy here: these are the pack/unpack contracts on either side of the hot
path (QueryParser.cpp:1146-1152,1347; DV-Kernel.cu:4268-4276; CPUfunctions.cpp:1093-1115).
"""
from __future__ import annotations

import numpy as np

ANSWER_NO_HIT = 0xFFFFFFFD       # DV-Kernel.cu:4474
ANSWER_OVERFLOW = 0xFFFFFFFE     # DV-Kernel.cu:4481
ANSWER_EMPTY = 0xFFFFFFFF        # DV-Kernel.cu:4278
ANSWER_OFFSET_LENGTH = 24        # DV-Kernel.h:30

NUM_CASES = {0: 1, 1: 2, 2: 4, 3: 6, 4: 10}                    # definitions.h:116-120
SA_RANGES_ROUND1 = {0: 2, 1: 4, 2: 4, 3: 2, 4: 1}              # definitions.h:47-59
SA_RANGES_ROUND2 = {0: 16, 1: 512, 2: 32, 3: 16, 4: 16}        # definitions.h:61-72


def ceil32(n: int) -> int:
    return (n + 31) // 32 * 32


def word_per_query(max_read_length: int) -> int:
    """definitions.h:444-454: next power of two of the read length, /16."""
    p = 1
    while p < max_read_length:
        p <<= 1
    return max(p // 16, 1)


def pack_queries(reads: np.ndarray, lengths: np.ndarray, wpq: int) -> np.ndarray:
    """reads: uint8 [N, Lmax] base codes 0..3 -> uint32[ceil32(N)*wpq] in the
    32-read word-interleaved layout; base i in bits 2*(i%16) of word i/16."""
    n, lmax = reads.shape
    assert lmax <= wpq * 16
    padded = np.zeros((ceil32(n), wpq * 16), dtype=np.uint32)
    mask = np.arange(lmax)[None, :] < lengths[:, None]
    padded[:n, :lmax] = np.where(mask, reads, 0)
    words = (padded.reshape(ceil32(n), wpq, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=2, dtype=np.uint64)
    words = words.astype(np.uint32)                       # [ceil32, wpq]
    out = words.reshape(-1, 32, wpq).transpose(0, 2, 1)   # [group, word, lane]
    return np.ascontiguousarray(out).reshape(-1)


def unpack_queries(q: np.ndarray, n: int, wpq: int) -> np.ndarray:
    """inverse of pack_queries -> uint8 [n, wpq*16]"""
    words = q.reshape(-1, wpq, 32).transpose(0, 2, 1).reshape(-1, wpq)[:n]
    bases = (words[:, :, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3
    return bases.reshape(n, wpq * 16).astype(np.uint8)


def answers_view(ans: np.ndarray, n: int, wpa: int) -> np.ndarray:
    """uint32[ceil32(n)*wpa] interleaved -> [n, wpa] row per read"""
    return ans.reshape(-1, wpa, 32).transpose(0, 2, 1).reshape(-1, wpa)[:n]


def decode_answer_row(row: np.ndarray):
    """-> (status, [(saL, saR, strand, mismatches), ...]) following
    CPUfunctions.cpp:1260-1283 (status 'ok' | 'none' | 'overflow')."""
    w0 = int(row[0])
    if w0 == ANSWER_NO_HIT:
        return "none", []
    status = "overflow" if w0 > ANSWER_NO_HIT else "ok"
    out = []
    for s in range(len(row) // 2):
        l, w = int(row[2 * s]), int(row[2 * s + 1])
        if s == 0 and status == "overflow":
            continue
        if l == ANSWER_EMPTY and w == ANSWER_EMPTY:
            break
        out.append((l, l + (w & ((1 << ANSWER_OFFSET_LENGTH) - 1)), (w >> 27) & 1, (w >> 24) & 7))
    return status, out


# ---------------------------------------------------------------------------
# DP batch formats (DV-DPfunctions.cu:55-59,1469-1524)
# ---------------------------------------------------------------------------

def dp_words(max_length: int) -> int:
    """MC_CeilDivide16 (DV-DPfunctions.h:59)"""
    return (max_length + 15) >> 4


def pack_dp_sequences(seqs: np.ndarray, max_length: int) -> np.ndarray:
    """seqs uint8 [B, <=max_length-1...] base codes, 0-based -> uint32[ceil32(B)*dp_words(max_length)]:
    base i (1-based!) in bits 2*(15-(i&15)) of word i>>4, slot 0 unused, 32-interleaved."""
    b, l = seqs.shape
    nw = dp_words(max_length)
    assert l + 1 <= nw * 16, "1-based packing needs length+1 <= 16*words"
    padded = np.zeros((ceil32(b), nw * 16), dtype=np.uint32)
    padded[:b, 1:l + 1] = seqs
    shifts = (2 * (15 - np.arange(16, dtype=np.uint32)))
    words = (padded.reshape(ceil32(b), nw, 16) << shifts).sum(axis=2, dtype=np.uint64).astype(np.uint32)
    return np.ascontiguousarray(words.reshape(-1, 32, nw).transpose(0, 2, 1)).reshape(-1)


def decode_pattern(pat: np.ndarray) -> str:
    """GPU pattern bytes -> CIGAR-like string in read order (the decode loop of
    SingleDP_Space::algnmtCPUThread, DV-DPfunctions.cu:1700-1716: 'V',n repeats the
    previous op n-1 more times; emitted right-to-left, so reverse at the end)."""
    ops = []
    i = 0
    last = None
    while i < len(pat) and pat[i] != 0:
        ch = chr(pat[i])
        if ch == 'V':
            cnt = int(pat[i + 1])
            ops.extend([last] * (cnt - 1) if cnt >= 1 else [])
            if cnt == 0 and ops:
                ops.pop()
            i += 2
        else:
            ops.append(ch)
            last = ch
            i += 1
    return "".join(reversed(ops))
