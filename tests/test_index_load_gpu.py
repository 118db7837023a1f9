"""s3_index_load: the index files soap3-dp-builder + BGS-Build write (the reference's own builders, oracle/_ref, when they are
there; else the same formats written from our builder's arrays, which the CPU tier pins bit for bit against those builders)
mapped and uploaded straight from disk == the index uploaded from arrays."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import ROOT, fmindex, formats, run_reference_builders, write_reference_files
from soap3dp_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_reference_builders", [False, True])
def test_index_load_equals_upload_from_arrays(use_reference_builders):
    G = synth.random_genome(250_003, seed=77)
    idx = fmindex.build_index(G, keep_sa=True)
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    with tempfile.TemporaryDirectory() as tmp:
        if use_reference_builders:
            if not os.path.exists(os.path.join(ref_dir, "soap3-dp-builder")):
                pytest.skip("oracle/_ref builders not shipped")
            prefix = run_reference_builders(G, tmp)
        else:
            prefix = os.path.join(tmp, "ours.index")
            write_reference_files(prefix, idx)
        g_files = api.index_load(prefix, with_text=True, device=0)
    g_mem = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    try:
        n, L, k = 4000, 100, 2
        rs = synth.simulate_single_end(G, n, L, seed=3, sub_rate=0.015)
        lens = np.zeros(formats.ceil32(n), np.uint32)
        lens[:n] = L
        wpq = formats.word_per_query(L)
        q = formats.pack_queries(rs.reads.numpy(), lens[:n], wpq)
        a = api.perform_round1_alignment(g_files, q, lens, n, wpq, k)
        b = api.perform_round1_alignment(g_mem, q, lens, n, wpq, k)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        # the text side: occurrences through the single-end chain (suffix array, packed text, check-and-extend)
        outs = []
        for g in (g_files, g_mem):
            al = api.SingleAligner(g, n, num_mismatch=k)
            outs.append(al.align(q, lens, n, wpq))
            al.free()
        for key in ("occ_offsets", "positions", "occ_flags", "read_flags"):
            assert np.array_equal(outs[0][key], outs[1][key]), key
        assert len(outs[0]["positions"]) > n // 2
    finally:
        api.GPUINDEXFree(g_files)
        api.GPUINDEXFree(g_mem)


def test_index_load_rejects_bad_files():
    with tempfile.TemporaryDirectory() as tmp:
        with pytest.raises(api.S3Error):
            api.index_load(os.path.join(tmp, "missing.index"))
        G = synth.random_genome(50_000, seed=5)
        idx = fmindex.build_index(G, keep_sa=True)
        prefix = os.path.join(tmp, "x.index")
        write_reference_files(prefix, idx)
        raw = np.fromfile(prefix + ".rev.fmv.gpu", np.uint32)
        raw[0] ^= 1                                                   # header of one file no longer agrees
        raw.tofile(prefix + ".rev.fmv.gpu")
        with pytest.raises(api.S3Error):
            api.index_load(prefix)
        raw[0] ^= 1
        raw.tofile(prefix + ".rev.fmv.gpu")
        sa = np.fromfile(prefix + ".sa", np.uint32)
        sa[5] = 4                                                     # a sampled suffix array
        sa.tofile(prefix + ".sa")
        with pytest.raises(api.S3Error):
            api.index_load(prefix, with_text=True)
        g = api.index_load(prefix, with_text=False)                   # the search-only index still loads
        api.GPUINDEXFree(g)
