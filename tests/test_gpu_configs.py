"""BASELINE.json configs[0] at its full size on the device -- the 10 Mb, four-chromosome genome and 100 000
single-end reads of 100 bases that bench.py's `configs` block times -- against the C oracle and against the
unmodified reference's SingleEndMapping (oracle/_ref/libwaltref.so) on EVERY read, through the same
device-resident call the bench times (walt_engine_map_se_device); and a two-pair slice of configs[3]'s
shape through the paired-end device call.  The index is built on the device and exported to the host for
the checkers, as in bench.py."""
import os
import sys
import types

import numpy as np
import pytest

import refio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _args(kind):
    return types.SimpleNamespace(group_width=8, table_depth=0, workload=kind, genome_mb=0.0, reads=0, read_len=0)


def test_configs0_full_size_every_read_vs_oracle_and_reference():
    import torch
    import bench
    if not refio.have_reference():
        pytest.skip("oracle/_ref was not built")
    gmb, n, rl = bench.FULL_SIZE["se_small"]
    assert (gmb, n, rl) == (10.0, 100_000, 100)
    wl = bench.Workload(_args("se_small"), 0, 0, kind="se_small", genome_mb=gmb, reads=n, read_len=rl)
    hidx = None
    try:
        assert len(wl.lengths) == 4
        stream = torch.cuda.current_stream()
        wl.device_step(stream.cuda_stream)
        torch.cuda.synchronize()
        assert wl.launches_per_step() >= 1
        got = wl.d_out.cpu().numpy().view(refio.BEST_DT)
        hidx = wl.host_index()
        ctr, obest = bench.oracle_counters(wl, hidx, n)
        for f in ("genome_pos", "times", "mismatch", "strand"):
            bad = np.nonzero(got[f] != obest[f])[0]
            assert bad.size == 0, (f, bad[:5], got[bad[:5]], obest[bad[:5]])
        _, rres = bench.reference_pass(wl, hidx, n, os.cpu_count() or 1)
        assert bench.parity_vs_reference(wl, n, rres) == 0
        assert 0.5 < float((got["times"] == 1).mean()) <= 1.0      # the reads do map
        assert ctr["n_lookups"] > n
    finally:
        if hidx is not None:
            bench.free_host_indexes(hidx)
        wl.close()


def test_configs3_shape_timed_paired_end_result_vs_reference():
    """configs[3]'s call (2 x 150 bp, -k 50 -L 1000) on a 20 Mb genome: the walt_pe_result records the timed
    device-resident call leaves behind against MergePairedEndResults on the reference's drained heaps."""
    import torch
    import bench
    if not refio.have_reference():
        pytest.skip("oracle/_ref was not built")
    n = 50_000
    wl = bench.Workload(_args("pe"), 0, 0, kind="pe", genome_mb=20.0, reads=n, read_len=150)
    hidx = None
    try:
        wl.device_step(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        hidx = wl.host_index()
        _, rres = bench.reference_pass(wl, hidx, n, os.cpu_count() or 1)
        assert bench.parity_vs_reference(wl, n, rres) == 0
        got = wl.d_out.cpu().numpy().view(wl.out_dt)
        assert float((got["pair"]["best_times"] == 1).mean()) > 0.5
    finally:
        if hidx is not None:
            bench.free_host_indexes(hidx)
        wl.close()
