"""Parity tests proper: the sm_100a kernels, called through the C ABI (include/bsbolt_b200.h), against
the reference aligner -- committed golden SAM, a live run of oracle/_ref/bwa when it travelled to the box,
and size-independent properties. Bit-exact: every SAM field and tag must be identical."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, first_diff, strip_pg
from smart_inter import SMART_CASES, write_interleaved

pytestmark = pytest.mark.gpu
CASES = ['se100', 'se100_un', 'se50_clip', 'pe150', 'pe150_un', 'pe150_un_sp0', 'pe150_opts']


@pytest.fixture(scope='module')
def index(built, golden):
    from bsbolt_b200 import _native
    assert _native.lib().bsb_device_count() >= 1, 'no CUDA device: the product has no CPU fallback'
    ix = _native.Index(golden.idxbase, 0)
    yield ix
    ix.close()


def run_mem(index, argv, tmp_path, tag):
    from bsbolt_b200 import _native
    out, log = tmp_path / f'{tag}.sam', tmp_path / f'{tag}.log'
    with open(out, 'w') as fo, open(log, 'w') as fl:
        rc, stats = _native.mem_main(argv, index=index, out_fd=fo.fileno(), log_fd=fl.fileno())
    assert rc == 0, _native.last_error() + open(log).read()[-2000:]
    bs = {}
    for l in open(log):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            bs[k] = bs.get(k, 0) + int(v)
    return open(out).read(), bs, stats


@pytest.mark.parametrize('case', CASES)
def test_golden_sam_bit_exact(index, golden, tmp_path, case):
    sam, bs, stats = run_mem(index, golden.argv(case), tmp_path, case)
    mine, want = strip_pg(sam), golden.sam(case)
    assert mine == want, first_diff(want, mine)
    assert bs == golden.cases[case]['bsstat']
    assert stats['kernel_launches'] > 0 and stats['ms_kernels'] > 0


@pytest.mark.parametrize('case', sorted(SMART_CASES))
def test_smart_pairing_golden_sam_bit_exact(index, golden, tmp_path, case):
    """`-p`: interleaved input split by read name into a single-end and a paired call per batch (fastmap.c:38-57)."""
    import gzip
    import json
    man = json.load(open(os.path.join(GOLDEN, 'smart_golden.json')))[case]
    fq = write_interleaved(golden.dir, tmp_path / 'smart_inter.fq')
    argv = ['mem'] + golden.manifest['launcher_args'] + man['extra'] + [golden.idxbase, fq]
    sam, bs, stats = run_mem(index, argv, tmp_path, case)
    mine, want = strip_pg(sam), gzip.open(os.path.join(GOLDEN, case + '.sam.gz'), 'rt').read()
    assert mine == want, first_diff(want, mine)
    assert bs == man['bsstat']
    assert stats['kernel_launches'] > 0


@pytest.mark.parametrize('case', ['long_se', 'long_se_w30'])
def test_long_reads_golden_sam_bit_exact(index, golden, tmp_path, case):
    """reads of 690-1200 bp: mem_flt_chained_seeds / mem_seed_sw (k_seed_sw) between chaining and extension"""
    import gzip
    import json
    man = json.load(open(os.path.join(GOLDEN, 'long_golden.json')))[case]
    argv = ['mem'] + golden.manifest['launcher_args'] + man['extra'] + [golden.idxbase] + [os.path.join(golden.dir, f) for f in man['fq']]
    sam, bs, stats = run_mem(index, argv, tmp_path, case)
    mine, want = strip_pg(sam), gzip.open(os.path.join(GOLDEN, case + '.sam.gz'), 'rt').read()
    assert mine == want, first_diff(want, mine)
    assert bs == man['bsstat']


def test_batch_api_matches_mem_main(index, golden, tmp_path):
    """bsb_batch_create/align/sam (host buffers in, SAM text out) == the file-based run"""
    from bsbolt_b200 import _native
    c = golden.cases['pe150']

    def load(fq):
        rs, f = [], open(os.path.join(golden.dir, fq))
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().strip(); f.readline(); q = f.readline().strip()
            rs.append((h[1:].split()[0], s, q))
        return rs
    r1, r2 = load(c['fq'][0]), load(c['fq'][1])
    opt = ['mem'] + golden.manifest['launcher_args']
    text, st = _native.align_batch(index, opt, r1, r2, n_processed=0)
    hdr_lines = [l for l in golden.sam('pe150').split('\n') if l.startswith('@')]
    want_full = golden.sam('pe150')
    # one batch holding all reads == the reference run when its -K covers the whole file: re-run the file API with a huge -K
    argv = ['mem'] + golden.manifest['launcher_args'] + ['-K', '100000000', golden.idxbase] + [os.path.join(golden.dir, f) for f in c['fq']]
    sam, _, _ = run_mem(index, argv, tmp_path, 'bigk')
    body = ''.join(l + '\n' for l in sam.split('\n') if l and not l.startswith('@'))
    assert text == body, first_diff(body, text)
    assert st['total_reads'] == len(r1)


def test_live_reference_on_fresh_reads(index, golden, tmp_path):
    """Differential run against the compiled reference on reads that are NOT in the fixtures:
    seeded mutations of the golden reads (so clipping, rescue, XA and unmapped paths all move)."""
    bwa = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
    if not os.path.exists(bwa):
        pytest.skip('oracle/_ref/bwa did not travel to this box')
    import random
    rnd = random.Random(1234)
    fqs = []
    for k, src in enumerate(('pe150uc_1.fq', 'pe150uc_2.fq')):
        dst = tmp_path / f'mut_{k}.fq'
        with open(os.path.join(golden.dir, src)) as f, open(dst, 'w') as o:
            while True:
                h = f.readline()
                if not h:
                    break
                s = list(f.readline().strip()); p = f.readline(); q = f.readline()
                if rnd.random() < 0.3:
                    for _ in range(rnd.randint(1, 12)):
                        s[rnd.randrange(len(s))] = rnd.choice('ACGTN')
                if rnd.random() < 0.05:
                    cut = rnd.randrange(20, len(s))
                    s = s[:cut]; q = q.strip()[:cut] + '\n'
                o.write(h + ''.join(s) + '\n' + p + q)
        fqs.append(str(dst))
    for extra in (['-z', '-K', '120000'], ['-K', '500000', '-L', '3,3', '-T', '20'], ['-z', '-e', '0', '-K', '90000', '-U', '9']):
        argv = ['mem'] + golden.manifest['launcher_args'] + extra + [golden.idxbase] + fqs
        ref = subprocess.run([bwa] + argv, capture_output=True, text=True)
        assert ref.returncode == 0
        sam, _, _ = run_mem(index, argv, tmp_path, 'live')
        a, b = strip_pg(ref.stdout), strip_pg(sam)
        assert a == b, first_diff(a, b)


def test_round_trip_properties(index, golden, tmp_path):
    """Size-independent checks on the device output: records in input order, one primary per read,
    CIGAR query length == read length, NM consistent with MD+CIGAR, reads drawn error-free from the
    reference align back to where they came from."""
    import gzip
    import random
    fa = {}
    name = None
    for l in gzip.open(os.path.join(ROOT, 'tests', 'golden', 'genome.fa.gz'), 'rt'):
        if l.startswith('>'):
            name = l[1:].strip(); fa[name] = []
        else:
            fa[name].append(l.strip())
    fa = {k: ''.join(v).upper() for k, v in fa.items()}
    rnd = random.Random(7)
    fq = tmp_path / 'exact.fq'
    truth = []
    comp = str.maketrans('ACGT', 'TGCA')
    with open(fq, 'w') as o:
        i = 0
        while i < 4000:
            c = rnd.choice(['chr2', 'chr3', 'chr4', 'chr5'])
            p = rnd.randrange(0, len(fa[c]) - 120)
            s = fa[c][p:p + 120]
            if 'N' in s:
                continue
            watson = rnd.random() < 0.5
            r = s.replace('C', 'T') if watson else s.replace('G', 'A')   # fully unmethylated, directional
            if not watson:
                r = r[::-1].translate(comp)                                # crick reads are sequenced as revcomp
            o.write(f'@r{i}\n{r}\n+\n{"I" * 120}\n')
            truth.append((c, p + 1, watson))
            i += 1
    argv = ['mem'] + golden.manifest['launcher_args'] + ['-K', '10000000', golden.idxbase, str(fq)]
    sam, bs, _ = run_mem(index, argv, tmp_path, 'exact')
    recs = [l.split('\t') for l in sam.split('\n') if l and not l.startswith('@')]
    prim = [r for r in recs if not int(r[1]) & 0x900]
    assert [r[0] for r in prim] == [f'r{i}' for i in range(4000)]
    hit = 0
    for r, (c, p, watson) in zip(prim, truth):
        if r[2] == c and int(r[3]) == p:
            hit += 1
            assert r[5] == '120M' and 'NM:i:0' in r and 'MD:Z:120' in r[11:13][1]
            assert ('YS:Z:W_C2T' in r) == watson and ('YS:Z:C_C2T' in r) == (not watson)
        clen = 0
        n = ''
        for ch in r[5]:
            if ch.isdigit():
                n += ch
            else:
                if ch in 'MIS=X':
                    clen += int(n)
                n = ''
        assert r[5] == '*' or clen == len(r[9])
    assert hit >= 3900


def test_index_build_byte_identical(built, golden, tmp_path):
    """GPU index builder (bsb_index_build) vs the files the reference's `bwa index` wrote for the same genome."""
    import hashlib
    from bsbolt_b200 import index_db
    ref, ms = index_db.build_database(os.path.join(golden.dir, 'genome.fa'), str(tmp_path / 'db'))
    for ext in ('pac', 'opac', 'ann', 'amb', 'bwt', 'sa'):
        got = hashlib.md5(open(f'{ref}.{ext}', 'rb').read()).hexdigest()
        assert got == golden.manifest['md5'][f'db/BSB_ref.fa.{ext}'], f'.{ext} differs from the reference index'
    assert ms > 0


def test_multi_device_run_identical(index, golden, tmp_path):
    """bsb_mem_main_multi (one reader, batch b on device b mod G, output in input order) over every visible GPU -- or two
    resident copies of the index on GPU 0 when the box has one -- prints the single-device SAM and BSStat lines; SAM and
    BAM output; every device gets its share of the batches."""
    from bsbolt_b200 import _native
    n = _native.lib().bsb_device_count()
    devices = list(range(n)) if n > 1 else [0, 0]
    multi = _native.MultiIndex(golden.idxbase, devices)
    try:
        for case in ('pe150', 'pe150_un', 'se100'):
            out, log = tmp_path / f'{case}.sam', tmp_path / f'{case}.log'
            with open(out, 'w') as fo, open(log, 'w') as fl:
                rc, st = _native.mem_main_multi(golden.argv(case), multi, out_fd=fo.fileno(), log_fd=fl.fileno())
            assert rc == 0, _native.last_error()
            got = strip_pg(open(out).read())
            assert got == golden.sam(case), first_diff(golden.sam(case), got)
            bs = {}
            for l in open(log):
                if l.startswith('BSStat '):
                    k, v = l[7:].split(': ')
                    bs[k] = bs.get(k, 0) + int(v)
            assert bs == golden.cases[case]['bsstat']
            assert st['n_batches'] >= len(devices) and st['kernel_launches'] > 0
        import gzip
        bam = tmp_path / 'multi.bam'
        with open(tmp_path / 'bam.log', 'w') as fl:
            rc, st = _native.mem_main_multi_bam(golden.argv('pe150'), str(bam), multi, threads=2, log_fd=fl.fileno())
        assert rc == 0, _native.last_error()
        one = tmp_path / 'one.bam'
        with open(tmp_path / 'bam1.log', 'w') as fl:
            rc, _ = _native.mem_main_bam(golden.argv('pe150'), str(one), index=multi.parts[0], threads=2, log_fd=fl.fileno())
        assert rc == 0
        a, b = gzip.open(bam, 'rb').read(), gzip.open(one, 'rb').read()
        assert a[:4] == b'BAM\x01' and len(a) == len(b)
    finally:
        multi.close()


def test_cli_drop_in_single_and_two_devices(built, golden, tmp_path):
    """`python -m bsbolt_b200 Align ... -OS` (the reference's CLI surface) on one device and over two (-GPU 0,0: two resident
    copies of the index driven by ONE process, batches dealt round-robin) -- both must print the reference's SAM."""
    import sys
    db = os.path.dirname(golden.idxbase)
    c = golden.cases['pe150']
    fq = [os.path.join(golden.dir, f) for f in c['fq']]
    want = golden.sam('pe150')
    for gpu in ('0', '0,0'):
        cmd = [sys.executable, '-m', 'bsbolt_b200', 'Align', '-F1', fq[0], '-F2', fq[1], '-DB', db, '-OS', '-K', '200000', '-GPU', gpu]
        p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
        assert p.returncode == 0, p.stderr[-2000:]
        assert strip_pg(p.stdout) == want, first_diff(want, strip_pg(p.stdout))
        assert 'Total Reads: 1954' in p.stderr


def test_bam_output(built, golden, tmp_path):
    """-O writes <prefix>.bam: BGZF container, BAM magic, every SAM record present."""
    import gzip
    import struct
    import sys
    db = os.path.dirname(golden.idxbase)
    c = golden.cases['se100']
    cmd = [sys.executable, '-m', 'bsbolt_b200', 'Align', '-F1', os.path.join(golden.dir, c['fq'][0]), '-DB', db, '-O', str(tmp_path / 'out'), '-K', '100000']
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    raw = gzip.open(tmp_path / 'out.bam', 'rb').read()   # BGZF is a series of gzip members
    assert raw[:4] == b'BAM\x01'
    l_text, = struct.unpack('<i', raw[4:8])
    off = 8 + l_text
    n_ref, = struct.unpack('<i', raw[off:off + 4]); off += 4
    for _ in range(n_ref):
        l_name, = struct.unpack('<i', raw[off:off + 4]); off += 4 + l_name + 4
    n = 0
    while off < len(raw):
        bs, = struct.unpack('<i', raw[off:off + 4]); off += 4 + bs; n += 1
    assert n_ref == 6 and n == c['n_records']
