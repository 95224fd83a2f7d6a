import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree native pieces exist (CUDA lib, host extension, oracle)."""
    import __graft_entry__ as g

    need = [os.path.join(ROOT, "mbf_bam_b200", "libmbfbam_cuda.so"),
            os.path.join(ROOT, "oracle", "libmbf_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    import mbf_bam_b200

    return mbf_bam_b200


@pytest.fixture(scope="session")
def has_gpu(built):
    return built.device_count() > 0


def make_case(tmpdir, seed, n_reads=3000, **kw):
    """(path, contigs, records, intervals, gene_intervals)"""
    from synth import bamio, small

    write_kw = {k: kw.pop(k) for k in list(kw) if k in ("level", "split_records", "payload_limit", "strategy", "sync_every")}
    contigs, recs, iv, giv = small.simple_case(seed, n_reads=n_reads, **kw)
    path = os.path.join(str(tmpdir), "case_%d.bam" % seed)
    bamio.write_bam(path, contigs, recs, **write_kw)
    return path, contigs, recs, iv, giv
