"""CPU tier: host-side formats, C-ABI surface, sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, formats
from soap3dp_b200 import api, sharding


def test_query_pack_roundtrip_and_layout():
    rng = np.random.default_rng(1)
    n, L = 70, 100
    reads = rng.integers(0, 4, (n, L)).astype(np.uint8)
    lens = np.full(n, L, np.uint32)
    wpq = formats.word_per_query(L)
    assert wpq == 8                                    # nextpow2(100)/16, definitions.h:444-454
    q = formats.pack_queries(reads, lens, wpq)
    assert q.size == formats.ceil32(n) * wpq
    # word w of read r at (r/32*32)*W + w*32 + r%32; base i at bits 2*(i%16) of word i/16
    r, i = 37, 53
    word = q[(r // 32 * 32) * wpq + (i // 16) * 32 + r % 32]
    assert (word >> (2 * (i % 16))) & 3 == reads[r, i]
    assert np.array_equal(formats.unpack_queries(q, n, wpq)[:, :L], reads)


def test_dp_pack_is_one_based_msb_first():
    seqs = np.array([[1, 2, 3, 0, 1]], np.uint8)
    w = formats.pack_dp_sequences(seqs, 20)
    assert w.size == 32 * 2
    word0 = int(w[0])
    for i in range(1, 6):
        assert (word0 >> ((15 - i) << 1)) & 3 == seqs[0, i - 1]
    assert (word0 >> 30) & 3 == 0                      # slot 0 unused


def test_answer_decode():
    row = np.array([5, (3) | (1 << 24) | (1 << 27), 0xFFFFFFFF, 0xFFFFFFFF], np.uint32)
    assert formats.decode_answer_row(row) == ("ok", [(5, 8, 1, 1)])
    assert formats.decode_answer_row(np.array([0xFFFFFFFD, 0xFFFFFFFF], np.uint32)) == ("none", [])
    st, _ = formats.decode_answer_row(np.array([0xFFFFFFFE, 7, 9, 1], np.uint32))
    assert st == "overflow"


def test_pattern_decode():
    pat = np.frombuffer(b"SV\x03MMmIDDM\x00", np.uint8)
    assert formats.decode_pattern(pat) == "MDDImMMSSS"


def test_c_abi_exports_every_declared_symbol():
    """the shared library loads without a GPU and exports exactly what include/*.h declares"""
    hdr = open(os.path.join(ROOT, "include", "soap3dp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(s3_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    lib = ctypes.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"


def test_no_cpu_fallback_without_device():
    """on a box without a GPU the product path must fail loudly, not fall back"""
    lib = api.load_library()
    if lib.s3_device_count() > 0:
        pytest.skip("a GPU is present")
    z = np.zeros(64, np.uint32)
    out = ctypes.c_void_p()
    rc = lib.s3_index_upload(api._u32(z), api._u32(z), api._u32(z), api._u32(z), 2, 0, 0, 100, None, None, 0,
                             ctypes.byref(out))
    assert rc != 0 and b"no CPU fallback" in lib.s3_last_error()
    with pytest.raises(api.S3Error):
        api.SemiGlobalAligner(104, 162, 64)


def test_product_sources_never_touch_the_oracle():
    pkg = os.path.join(ROOT, "soap3-dp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "liboracle" not in src and "oracle/" not in src.replace("oracle/dp_oracle.c has", ""), f


def test_shard_ranges():
    r = sharding.shard_ranges(1000, 3)
    assert r[0][0] == 0 and r[-1][1] == 1000
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(b % 32 == 0 for b, _ in r)
    assert sharding.shard_ranges(10, 4) == [(0, 10), (10, 10), (10, 10), (10, 10)]
    assert sharding.shard_ranges(0, 2) == [(0, 0), (0, 0)]
    sizes = [e - b for b, e in sharding.shard_ranges(50_000_000, 8)]
    assert max(sizes) - min(sizes) <= 32


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["S3_ROOT"])
import _pkg; _pkg.load()
from soap3dp_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1000
b, e = sharding.shard_ranges(n, world)[rank]
local = np.arange(b, e, dtype=np.int64)[:, None] * np.array([1, 10])     # per-unit "result"
allr = sharding.gather_to_rank0(local)
if rank == 0:
    assert allr.shape == (n, 2) and (allr[:, 0] == np.arange(n)).all() and (allr[:, 1] == 10 * np.arange(n)).all()
    print("GLOO_OK")
else:
    assert allr is None
dist.destroy_process_group()
'''


def test_sharded_results_come_back_in_input_order_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, S3_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert "GLOO_OK" in out.stdout, out.stdout + out.stderr

