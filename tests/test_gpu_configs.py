"""BASELINE configs C1 and C3 on the REFERENCE's own test genome (tests/TestData/BSB_test.fa, committed as
tests/golden/c1/BSB_test.fa.gz; goldens by tests/golden/make_c1c3_golden.py from the reference itself):

  * the database is written by the product (`bsbolt_b200.index_db`: whole genome, `-MR` bed-masked, `-rrbs` MspI) and indexed by
    the GPU builder; every file must have the md5 of what the reference's `bsbolt Index` + `bwa index` wrote (and, when
    oracle/_ref/bwa travelled to the box, of a live `bwa index` of the same FASTA);
  * C1: the reference simulator's SE100 reads at depth 5 (`bsbolt Simulate -RL 100 -RD 5`, the PR1 set): 100 % of the
    SAM records identical to the reference aligner's (committed md5 + a live run), directional and `-UN`;
  * C3: SE50 reads of an in-silico MspI library (~120 reads per start site: the high-duplicate, short-seed regime) on
    the RRBS-masked database, and the same reads on the bed-masked database (where most of them have nowhere to go).
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess

import pytest

from conftest import GOLDEN, ROOT, first_diff, strip_pg

pytestmark = pytest.mark.gpu
BWA = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
MAN = json.load(open(os.path.join(GOLDEN, 'c1c3.json')))
EXTS = ('amb', 'ann', 'bwt', 'sa', 'pac', 'opac')


def md5(path):
    return hashlib.md5(open(path, 'rb').read()).hexdigest()


@pytest.fixture(scope='module')
def testdata(built, tmp_path_factory):
    from bsbolt_b200 import _native, index_db
    if _native.lib().bsb_device_count() < 1:
        pytest.fail('no CUDA device: the product has no CPU fallback')
    d = tmp_path_factory.mktemp('c1c3')
    fa = str(d / 'BSB_test.fa')
    with gzip.open(os.path.join(GOLDEN, 'c1', 'BSB_test.fa.gz'), 'rb') as i, open(fa, 'wb') as o:
        shutil.copyfileobj(i, o)
    assert md5(fa) == MAN['fasta_md5']
    bed = os.path.join(GOLDEN, 'c1', 'test_wgbs_masking.bed')
    dbs = {}
    dbs['wgbs'], _ = index_db.build_database(fa, str(d / 'db_wgbs'))
    dbs['masked'], _ = index_db.build_database(fa, str(d / 'db_masked'), mappable_regions=bed)
    dbs['rrbs'], _ = index_db.build_rrbs_database(fa, str(d / 'db_rrbs'))

    class T:
        dir = d
        fasta = fa
        db = dbs
    return T


@pytest.mark.parametrize('mode', ['wgbs', 'masked', 'rrbs'])
def test_database_files_identical_to_bsbolt_index(testdata, mode, tmp_path):
    ref = testdata.db[mode]
    want = MAN['db'][mode]
    assert md5(ref) == want['BSB_ref.fa']
    for ext in EXTS:
        assert md5(f'{ref}.{ext}') == want[f'BSB_ref.fa.{ext}'], f'{mode}: .{ext} differs from the reference index'
    if mode == 'rrbs':
        got = hashlib.md5(gzip.open(os.path.join(os.path.dirname(ref), 'mappable_regions.bed.gz'), 'rb').read()).hexdigest()
        assert got == want['mappable_regions.bed']
    if os.path.exists(BWA):   # and against the reference indexer run here on the same FASTA
        live = str(tmp_path / 'BSB_ref.fa')
        shutil.copy(ref, live)
        p = subprocess.run([BWA, 'index', '-a', 'bwtsw', live], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        for ext in EXTS:
            assert md5(f'{ref}.{ext}') == md5(f'{live}.{ext}'), f'{mode}: .{ext} differs from a live bwa index'


def align_and_compare(db, fqs, case, tmp_path):
    from bsbolt_b200 import _native
    argv = ['mem'] + MAN['launcher_args'] + case['extra'] + [db] + fqs
    ix = _native.Index(db, 0)
    out, log = tmp_path / 'mine.sam', tmp_path / 'mine.log'
    with open(out, 'w') as fo, open(log, 'w') as fl:
        rc, st = _native.mem_main(argv, index=ix, out_fd=fo.fileno(), log_fd=fl.fileno())
    ix.close()
    assert rc == 0, _native.last_error()
    mine = strip_pg(open(out).read())
    if os.path.exists(BWA):
        ref = subprocess.run([BWA] + argv, capture_output=True, text=True)
        assert ref.returncode == 0, ref.stderr[-2000:]
        assert strip_pg(ref.stdout) == mine, first_diff(strip_pg(ref.stdout), mine)
    assert sum(1 for l in mine.split('\n') if l and l[0] != '@') == case['n_records']
    assert hashlib.md5(mine.encode()).hexdigest() == case['sam_md5']
    bs = {}
    for l in open(log):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            bs[k] = bs.get(k, 0) + int(v)
    assert bs == case['bsstat']
    return st


@pytest.fixture(scope='module')
def c1_reads(testdata):
    fq = str(testdata.dir / 'se100_1.fq')
    with gzip.open(os.path.join(GOLDEN, 'c1', 'se100_1.fq.gz'), 'rb') as i, open(fq, 'wb') as o:
        shutil.copyfileobj(i, o)
    assert md5(fq) == MAN['c1']['fq_md5']
    return fq


def test_c1_se100_directional_all_records_identical(testdata, c1_reads, tmp_path):
    st = align_and_compare(testdata.db['wgbs'], [c1_reads], MAN['c1'], tmp_path)
    assert st['n_batches'] >= 4


def test_c1_se100_undirectional(testdata, c1_reads, tmp_path):
    align_and_compare(testdata.db['wgbs'], [c1_reads], MAN['c1_undirectional'], tmp_path)


@pytest.fixture(scope='module')
def c3_reads(testdata):
    from bsbolt_b200 import simulate
    names, seqs = simulate.read_fasta(testdata.fasta)
    fq = str(testdata.dir / 'rrbs50.fq')
    simulate.simulate_rrbs_reads(names, seqs, fq, **MAN['c3']['reads'])
    assert md5(fq) == MAN['c3']['fq_md5'], 'the seeded RRBS simulator no longer reproduces the reads the golden SAM was made from'
    return fq


def test_c3_rrbs_se50_on_mspi_masked_database(testdata, c3_reads, tmp_path):
    st = align_and_compare(testdata.db['rrbs'], [c3_reads], MAN['c3'], tmp_path)
    assert st['n_batches'] >= 5


def test_c3_reads_on_bed_masked_database(testdata, c3_reads, tmp_path):
    align_and_compare(testdata.db['masked'], [c3_reads], MAN['c3_on_masked_wgbs'], tmp_path)
