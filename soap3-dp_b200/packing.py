"""torch (CPU or CUDA) versions of the buffer packers in formats.py, used to build
benchmark-sized batches directly in HBM.  Same layouts, tested against the numpy
versions (tests/test_cpu_host_logic.py)."""
from __future__ import annotations

import torch


def _u32_to_i32(x: torch.Tensor) -> torch.Tensor:
    return torch.where(x >= (1 << 31), x - (1 << 32), x).to(torch.int32)


def pack_queries(reads: torch.Tensor, lengths: torch.Tensor, wpq: int) -> torch.Tensor:
    """reads uint8 [N, L] -> int32 (uint32 bit patterns) [ceil32(N)*wpq], 32-read word
    interleave, base i in bits 2*(i%16) of word i/16 (QueryParser.cpp:1146-1152)."""
    n, lmax = reads.shape
    dev = reads.device
    up = (n + 31) // 32 * 32
    out = torch.zeros(up // 32, wpq, 32, dtype=torch.int32, device=dev)
    shifts = (2 * torch.arange(16, device=dev)).to(torch.int64)
    step = 1 << 20                      # reads per chunk, multiple of 32
    ar = torch.arange(lmax, device=dev)
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        m = r1 - r0
        mup = (m + 31) // 32 * 32
        padded = torch.zeros(mup, wpq * 16, dtype=torch.int64, device=dev)
        seg = reads[r0:r1].to(torch.int64)
        seg = torch.where(ar[None, :] < lengths[r0:r1, None], seg, torch.zeros_like(seg))
        padded[:m, :lmax] = seg
        words = (padded.view(mup, wpq, 16) << shifts).sum(dim=2)
        out[r0 // 32:(r0 + mup) // 32] = _u32_to_i32(words).view(-1, 32, wpq).transpose(1, 2)
    return out.reshape(-1)


def pack_dp_sequences(seqs: torch.Tensor, max_length: int) -> torch.Tensor:
    """seqs uint8 [B, l] (0-based) -> int32 [ceil32(B)*ceil(max_length/16)]: 1-based, base i
    in bits 2*(15-(i&15)) of word i>>4, 32-interleaved (DV-DPfunctions.cu:57-59)."""
    b, l = seqs.shape
    dev = seqs.device
    nw = (max_length + 15) >> 4
    assert l + 1 <= nw * 16
    up = (b + 31) // 32 * 32
    out = torch.zeros(up // 32, nw, 32, dtype=torch.int32, device=dev)
    shifts = (2 * (15 - torch.arange(16, device=dev))).to(torch.int64)
    step = 1 << 17
    for r0 in range(0, b, step):
        r1 = min(b, r0 + step)
        m = r1 - r0
        mup = (m + 31) // 32 * 32
        padded = torch.zeros(mup, nw * 16, dtype=torch.int64, device=dev)
        padded[:m, 1:l + 1] = seqs[r0:r1].to(torch.int64)
        words = (padded.view(mup, nw, 16) << shifts).sum(dim=2)
        out[r0 // 32:(r0 + mup) // 32] = _u32_to_i32(words).view(-1, 32, nw).transpose(1, 2)
    return out.reshape(-1)


def answers_status(answers: torch.Tensor, n: int, wpa: int) -> torch.Tensor:
    """word 0 of every read's slot as int64 [n] (0xFFFFFFFD none, 0xFFFFFFFE overflow)"""
    w0 = answers.view(-1, wpa, 32)[:, 0, :].reshape(-1)[:n].to(torch.int64)
    return torch.where(w0 < 0, w0 + (1 << 32), w0)
