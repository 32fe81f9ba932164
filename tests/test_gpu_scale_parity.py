"""Parity at an out-of-cache index size: a 48 Mb synthetic genome (192 M-symbol BWT, 96 MB + SA -- far beyond
what the golden 2 Mb fixtures exercise, so occ blocks, SA and pac fetches really come from HBM), reads from the
seeded simulator (substitutions, indels, 5 % heavily corrupted mates so that mate-rescue Smith-Waterman runs),
aligned by the sm_100a path through the C ABI and by the compiled reference (oracle/_ref/bwa) on the same
files with the same argv. Every SAM record must be identical. Covers BASELINE configs C2 (PE150 directional),
C3-like (SE50), C5 (undirectional PE150 + rescue), and the reference-layout sampled suffix array
(the path a >= 2^32-symbol index such as C4 uses) next to the dense one."""
import os
import subprocess

import pytest

from conftest import ROOT, first_diff, strip_pg

pytestmark = pytest.mark.gpu
BWA = os.path.join(ROOT, 'oracle', '_ref', 'bwa')


@pytest.fixture(scope='module')
def big(built, golden, tmp_path_factory):
    from bsbolt_b200 import _native, index_db, simulate
    if _native.lib().bsb_device_count() < 1:
        pytest.fail('no CUDA device: the product has no CPU fallback')
    if not os.path.exists(BWA):
        pytest.skip('oracle/_ref/bwa did not travel to this box')
    d = tmp_path_factory.mktemp('big')
    fa = str(d / 'genome.fa')
    names, contigs = simulate.make_genome(fa, [8000000] * 6, seed=4242, n_dups=12)
    index_db.build_database(fa, str(d / 'db'), device=0)

    class B:
        dir = d
        db = str(d / 'db' / 'BSB_ref.fa')
        args = golden.manifest['launcher_args']
    B.names, B.contigs = names, contigs
    return B


def both(big, extra, fqs, tmp_path, env=None):
    from bsbolt_b200 import _native
    argv = ['mem'] + big.args + extra + [big.db] + fqs
    ref = subprocess.run([BWA] + argv + [], capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr[-2000:]
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k); os.environ[k] = v
    try:
        ix = _native.Index(big.db, 0)
        out, log = tmp_path / 'mine.sam', tmp_path / 'mine.log'
        with open(out, 'w') as fo, open(log, 'w') as fl:
            rc, st = _native.mem_main(argv, index=ix, out_fd=fo.fileno(), log_fd=fl.fileno())
        ix.close()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    assert rc == 0, _native.last_error()
    a, b = strip_pg(ref.stdout), strip_pg(open(out).read())
    n = sum(1 for l in a.split('\n') if l and not l.startswith('@'))
    assert n > 0
    assert a == b, first_diff(a, b)
    return n, st


def test_48mb_index_identical_to_bwa_index(big, tmp_path):
    """The out-of-cache index every test of this module (and, at 250 Mb, the benchmark) aligns against is built by the
    product's GPU builder; here the REFERENCE indexer (`oracle/_ref/bwa index -a bwtsw`, bwtindex.c:256-321) builds the
    same database FASTA on the host and all six files must be byte-identical -- so a fault of the builder that both
    aligners would digest the same way (N-run lrand48 fill, long contigs, SA sampling) cannot hide behind matching SAM."""
    import hashlib
    import shutil
    live = str(tmp_path / 'BSB_ref.fa')
    shutil.copy(big.db, live)
    p = subprocess.run([BWA, 'index', '-a', 'bwtsw', live], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    for ext in ('amb', 'ann', 'pac', 'opac', 'bwt', 'sa'):
        a = hashlib.md5(open(f'{big.db}.{ext}', 'rb').read()).hexdigest()
        b = hashlib.md5(open(f'{live}.{ext}', 'rb').read()).hexdigest()
        assert a == b, f'.{ext} of the GPU-built 48 Mb index differs from bwa index'


def test_c2_pe150_directional_with_rescue(big, tmp_path):
    from bsbolt_b200 import simulate
    fqs, n = simulate.simulate_reads(big.names, big.contigs, str(tmp_path / 'pe'), 40000, seed=11, corrupt_frac=0.05)
    recs, st = both(big, ['-K', '6000000', '-t', '16'], fqs, tmp_path)
    assert recs >= 2 * n and st['n_batches'] >= 2


def test_c2_sampled_suffix_array_layout(big, tmp_path):
    """The configuration a >= 2^32-symbol index (C4) runs with: BSB_SAMPLED_SA keeps the reference's SA layout (every
    32nd rank, u64, LF-walk on the device) and BSB_REF_BLOCKS seeds over the reference's 64-byte occ blocks with 64-bit
    counts instead of the sector-sized 32-bit ones."""
    from bsbolt_b200 import simulate
    fqs, n = simulate.simulate_reads(big.names, big.contigs, str(tmp_path / 'pe'), 15000, seed=12, corrupt_frac=0.02)
    both(big, ['-K', '100000000', '-t', '16'], fqs, tmp_path, env={'BSB_SAMPLED_SA': '1', 'BSB_REF_BLOCKS': '1'})
    # the same wide configuration with the 40-bit dense suffix array (5 bytes per rank) a human-scale index gets when it fits
    both(big, ['-K', '100000000', '-t', '16'], fqs, tmp_path, env={'BSB_DENSE_SA40': '1', 'BSB_REF_BLOCKS': '1'})


def test_c5_undirectional_pe150(big, tmp_path):
    from bsbolt_b200 import simulate
    fqs, n = simulate.simulate_reads(big.names, big.contigs, str(tmp_path / 'un'), 20000, seed=13, undirectional=True, corrupt_frac=0.05)
    both(big, ['-z', '-K', '4000000', '-t', '16'], fqs, tmp_path)
    both(big, ['-z', '-e', '0', '-K', '4000000', '-t', '16'], fqs, tmp_path)   # every read under both conversion patterns


def test_c3_se50_short_reads(big, tmp_path):
    from bsbolt_b200 import simulate
    fqs, n = simulate.simulate_reads(big.names, big.contigs, str(tmp_path / 'se'), 40000, read_len=50, paired=False, seed=14)
    both(big, ['-K', '1000000', '-t', '16'], fqs, tmp_path)
