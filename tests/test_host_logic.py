"""Host-side mirror of the reference interface: CLI flags, argv construction, exceptions, BAM encoder,
shard merge. No GPU."""
import gzip
import io
import os
import struct
import sys

import pytest

from conftest import ROOT

REF = '/root/reference'


def test_align_flags_match_reference_parser():
    from bsbolt_b200.Utils.Parser import ALIGN_FLAGS, parser
    flags = [f for f, _ in ALIGN_FLAGS]
    expected = ['-F1', '-F2', '-UN', '-O', '-OS', '-DB', '-CP', '-CT', '-SP', '-t', '-k', '-w', '-d', '-r', '-y', '-c', '-D', '-W', '-m',
                '-S', '-P', '-A', '-B', '-INDEL', '-E', '-L', '-U', '-p', '-R', '-H', '-j', '-T', '-XA', '-DR', '-M', '-I', '-OT']
    assert flags == expected
    if os.path.exists(os.path.join(REF, 'bsbolt', 'Utils', 'Parser.py')):
        src = open(os.path.join(REF, 'bsbolt', 'Utils', 'Parser.py')).read()
        import re
        ref_flags = re.findall(r"align_parser\.add_argument\('(-\w+)'", src)
        assert ref_flags == expected
    a = parser.parse_args(['Align', '-F1', 'a.fq', '-DB', 'db', '-OS'])
    assert (a.t, a.k, a.T, a.L, a.XA, a.DR, a.INDEL, a.U) == (1, 19, 10, '30,30', '100,200', 0.95, '6,6', 17)


def test_index_flags_match_reference_parser():
    from bsbolt_b200.Utils.Parser import INDEX_FLAGS, parser
    flags = [f for f, _ in INDEX_FLAGS]
    expected = ['-G', '-DB', '-B', '-MR', '-IA', '-rrbs', '-rrbs-cut-format', '-rrbs-lower', '-rrbs-upper']
    assert flags == expected
    if os.path.exists(os.path.join(REF, 'bsbolt', 'Utils', 'Parser.py')):
        src = open(os.path.join(REF, 'bsbolt', 'Utils', 'Parser.py')).read()
        import re
        assert re.findall(r"index_parser\.add_argument\('(-[\w-]+)'", src) == expected
    a = parser.parse_args(['Index', '-G', 'g.fa', '-DB', 'db', '-rrbs'])
    assert (a.rrbs, a.rrbs_cut_format, a.rrbs_lower, a.rrbs_upper, a.MR, a.IA, a.B) == (True, 'C-CGG', 40, 500, None, False, 10000000)


def test_argv_equals_reference_launcher(tmp_path):
    """build_alignment_command() produces the argv the reference's launch_alignment() would exec."""
    from bsbolt_b200.Utils.Launcher import build_alignment_command
    from bsbolt_b200.Utils.Parser import parser
    db = tmp_path / 'db'
    db.mkdir()
    (db / 'BSB_ref.fa').write_text('>x\nA\n'); (db / 'BSB_ref.fa.opac').write_text('')
    (tmp_path / 'r1.fq').write_text(''); (tmp_path / 'r2.fq').write_text('')
    a = parser.parse_args(['Align', '-F1', str(tmp_path / 'r1.fq'), '-F2', str(tmp_path / 'r2.fq'), '-DB', str(db), '-OS', '-UN', '-t', '4', '-M'])
    cmd = build_alignment_command(a)
    assert cmd[1:] == ['mem', '-Y', '-z', '-M', '-A', '1', '-B', '4', '-D', '0.5', '-E', '1,1', '-L', '30,30', '-T', '10', '-U', '17', '-W', '0',
                       '-c', '500', '-d', '100', '-k', '19', '-m', '50', '-r', '1.5', '-t', '4', '-w', '100', '-y', '20', '-O', '6,6',
                       '-h', '100,200', '-e', '0.1', '-l', '0.5', '-n', '5', '-Z', '0.95', f'{db}/BSB_ref.fa', str(tmp_path / 'r1.fq'), str(tmp_path / 'r2.fq')]


def test_api_surface():
    from bsbolt_b200.Align.AlignReads import (AlignmentCompressionError, BisulfiteAlignmentAndProcessing,
                                              BisulfiteAlignmentError)
    b = BisulfiteAlignmentAndProcessing(['bwa', 'mem', 'x'], output='o', output_threads=2, output_to_stdout=True)
    assert sorted(b.mapping_statistics) == sorted(['TotalReads', 'TotalAlignments', 'BSAmbiguous', 'C_C2T', 'C_G2A', 'W_C2T', 'W_G2A', 'Unaligned'])
    assert issubclass(BisulfiteAlignmentError, Exception) and issubclass(AlignmentCompressionError, Exception)
    with pytest.raises(BisulfiteAlignmentError):
        BisulfiteAlignmentAndProcessing(['bwa', 'index'], output_to_stdout=True).align_reads()


def test_bam_encoder_round_trip(built, golden, tmp_path):
    from bsbolt_b200.Utils.BamOutput import sam_file_to_bam
    sam = golden.sam('pe150')
    (tmp_path / 'x.sam').write_text(sam)
    out = tmp_path / 'x.bam'
    assert sam_file_to_bam(str(tmp_path / 'x.sam'), str(out), threads=2) == golden.cases['pe150']['n_records']
    raw = gzip.open(out, 'rb').read()
    assert raw[:4] == b'BAM\x01'
    l_text, = struct.unpack('<i', raw[4:8])
    assert raw[8:8 + l_text].decode() == ''.join(l + '\n' for l in sam.split('\n') if l.startswith('@'))
    off = 8 + l_text
    n_ref, = struct.unpack('<i', raw[off:off + 4]); off += 4
    names = []
    for _ in range(n_ref):
        l_name, = struct.unpack('<i', raw[off:off + 4]); off += 4
        names.append(raw[off:off + l_name - 1].decode()); off += l_name + 4
    recs = [l.split('\t') for l in sam.split('\n') if l and not l.startswith('@')]
    for r in recs:
        bs, = struct.unpack('<i', raw[off:off + 4])
        rid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, nrid, npos, tlen = struct.unpack('<iiBBHHHIiii', raw[off + 4:off + 36])
        assert (names[rid] if rid >= 0 else '*') == r[2] and pos + 1 == int(r[3]) and mapq == int(r[4]) and flag == int(r[1])
        assert raw[off + 36:off + 36 + l_rn - 1].decode() == r[0] and tlen == int(r[8]) and l_seq == (0 if r[9] == '*' else len(r[9]))
        off += 4 + bs
    assert off == len(raw)
    assert open(out, 'rb').read()[-28:] == bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')


def test_merge_shards(tmp_path):
    from bsbolt_b200.shard import merge_shards
    (tmp_path / 'a.sam').write_bytes(b'HDR\nb0\nb2\n'); (tmp_path / 'a.idx').write_text('-1\t0\t4\n0\t4\t3\n2\t7\t3\n')
    (tmp_path / 'b.sam').write_bytes(b'b1\nb3\n'); (tmp_path / 'b.idx').write_text('1\t0\t3\n3\t3\t3\n')
    out = io.BytesIO()
    assert merge_shards([str(tmp_path / 'a.sam'), str(tmp_path / 'b.sam')], [str(tmp_path / 'a.idx'), str(tmp_path / 'b.idx')], out) == 5
    assert out.getvalue() == b'HDR\nb0\nb1\nb2\nb3\n'


def test_reader_variants_give_identical_batches(built, golden, tmp_path):
    """The fast plain-FASTQ parser and the general kseq-style parser (gzip, CRLF, missing final newline, a record the
    fast parser does not recognise in the middle of the file) must cut identical reads and batches."""
    import gzip
    import subprocess
    hostsim = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim')
    src = open(os.path.join(golden.dir, golden.cases['se100']['fq'][0])).read().split('\n')
    recs = [r for r in (src[i:i + 4] for i in range(0, 4 * 3000, 4)) if len(r) == 4 and r[0]]
    plain = ''.join('\n'.join(r) + '\n' for r in recs)
    variants = {'plain.fq': plain.encode(), 'nonl.fq': plain[:-1].encode(), 'crlf.fq': plain.replace('\n', '\r\n').encode()}
    # the middle record with its sequence and quality folded over two lines each: only the general parser reads that
    folded = []
    for k, r in enumerate(recs):
        if k == len(recs) // 2:
            h = len(r[1]) // 2
            folded.append('\n'.join([r[0], r[1][:h], r[1][h:], r[2], r[3][:h], r[3][h:]]) + '\n')
        else:
            folded.append('\n'.join(r) + '\n')
    variants['folded.fq'] = ''.join(folded).encode()
    outs = {}
    for name, data in variants.items():
        (tmp_path / name).write_bytes(data)
    with gzip.open(tmp_path / 'plain.fq.gz', 'wb') as f:
        f.write(plain.encode())
    for name in list(variants) + ['plain.fq.gz']:
        argv = [hostsim, 'mem'] + golden.manifest['launcher_args'] + ['-K', '70000', golden.idxbase, str(tmp_path / name)]
        p = subprocess.run(argv, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-1500:]
        outs[name] = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
    env = dict(os.environ, BSB_SLOW_READER='1')
    p = subprocess.run([hostsim, 'mem'] + golden.manifest['launcher_args'] + ['-K', '70000', golden.idxbase, str(tmp_path / 'plain.fq')],
                       capture_output=True, text=True, env=env)
    want = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
    assert want.count('\n') > 1000
    for name, got in outs.items():
        assert got == want, name
    # the multi-threaded cutter (pieces of the file parsed speculatively, chained in file order): tiny pieces so that this
    # small file spans hundreds of them, piece boundaries on every kind of line; qualities that begin with '@' or '+' (a
    # quality line must never be taken for a header), and the variants above, where the serial / general parser takes over
    adv = ''.join('\n'.join([r[0], r[1], r[2], ('@' if k % 3 == 0 else '+' if k % 3 == 1 else 'I') + r[3][1:]]) + '\n' for k, r in enumerate(recs) if len(r) == 4 and r[0])
    (tmp_path / 'adv.fq').write_text(adv)
    p = subprocess.run([hostsim, 'mem'] + golden.manifest['launcher_args'] + ['-K', '70000', golden.idxbase, str(tmp_path / 'adv.fq')], capture_output=True, text=True, env=env)
    want_adv = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
    assert want_adv.count('\n') > 1000 and want_adv != want
    for piece, threads in (('97', '3'), ('1000', '2'), ('4099', '5'), ('100000', '4')):
        penv = dict(os.environ, BSB_READ_PIECE=piece, BSB_PARSE_THREADS=threads, BSB_DEBUG_READER='1')
        for name, expect in [(n, want) for n in variants] + [('adv.fq', want_adv)]:
            p = subprocess.run([hostsim, 'mem'] + golden.manifest['launcher_args'] + ['-K', '70000', golden.idxbase, str(tmp_path / name)],
                               capture_output=True, text=True, env=penv)
            assert p.returncode == 0, p.stderr[-1500:]
            got = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
            assert got == expect, (name, piece, threads)
            note = [l for l in p.stderr.split('\n') if l.startswith('[D::reader]')]
            assert note, 'the multi-threaded cutter did not run'
            if name in ('plain.fq', 'adv.fq'):
                assert 'whole file cut in parallel' in note[0], note
            else:
                assert 'serial parser takes over' in note[0], note


def test_database_fasta_writers_match_bsbolt_index(tmp_path):
    """`bsbolt Index` front half (whole genome, -MR bed mask, -rrbs MspI mask) on the reference's own test genome: the
    FASTA handed to the indexer and mappable_regions.bed must be byte-identical to what the reference wrote
    (md5s recorded by tests/golden/make_c1c3_golden.py from bsbolt/Index/{WholeGenomeIndex,RRBSIndex}.py)."""
    import gzip
    import hashlib
    import json
    import shutil
    from bsbolt_b200 import index_db
    g = os.path.join(ROOT, 'tests', 'golden')
    man = json.load(open(os.path.join(g, 'c1c3.json')))
    fa = str(tmp_path / 'BSB_test.fa')
    with gzip.open(os.path.join(g, 'c1', 'BSB_test.fa.gz'), 'rb') as i, open(fa, 'wb') as o:
        shutil.copyfileobj(i, o)
    md5 = lambda p: hashlib.md5(open(p, 'rb').read()).hexdigest()
    assert md5(index_db.write_database_fasta(fa, str(tmp_path / 'w'))) == man['db']['wgbs']['BSB_ref.fa']
    assert md5(index_db.write_database_fasta(fa, str(tmp_path / 'm'), mappable_regions=os.path.join(g, 'c1', 'test_wgbs_masking.bed'))) == man['db']['masked']['BSB_ref.fa']
    assert md5(index_db.write_rrbs_database_fasta(fa, str(tmp_path / 'r'))) == man['db']['rrbs']['BSB_ref.fa']
    assert hashlib.md5(gzip.open(tmp_path / 'r' / 'mappable_regions.bed.gz', 'rb').read()).hexdigest() == man['db']['rrbs']['mappable_regions.bed']
    # restriction-site table: IUPAC expansion and cut offsets (RRBSCutSites.py)
    assert index_db.restriction_sites('C-CGG') == {'CCGG': 1}
    assert index_db.restriction_sites('T-CGA,GR-CGYC') == {'TCGA': 1, 'GACGCC': 4, 'GGCGTC': 2, 'GGCGCC': 2, 'GACGTC': 2}   # values of ProcessCutSites
    assert index_db.restriction_sites('CCWGG') == {'CCAGG': 0, 'CCTGG': 0}
    # the one-step-per-position region walk of the reference's masking loop, on nested / out-of-order regions
    def ref_mask(seq, regions):
        it = iter(regions)
        start, end = next(it)
        out = []
        for pos, bp in enumerate(seq):
            if pos > end:
                try:
                    start, end = next(it)
                except StopIteration:
                    pass
            out.append(bp if start <= pos <= end else '-')
        return ''.join(out)
    import random
    rnd = random.Random(5)
    seq = ''.join(rnd.choice('ACGT') for _ in range(300))
    for _ in range(200):
        regs = sorted(((rnd.randrange(-5, 300), rnd.randrange(-5, 320)) for _ in range(rnd.randrange(1, 12))), key=lambda r: r[0])
        assert index_db.mask_outside(seq, regs) == ref_mask(seq, regs)


@pytest.mark.parametrize('case,n_dev', [('pe150', 2), ('pe150_un', 3), ('se100', 2)])
def test_multi_device_pipeline_equals_single_run(built, golden, case, n_dev):
    """bsb_mem_main_multi's host logic on CPU "devices" (tests/hostsim, HOSTSIM_DEVICES): one reader, batch b on device
    b mod G, several batches in flight, results in input order -- SAM and BSStat lines identical to the reference."""
    import subprocess
    hostsim = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim')
    env = dict(os.environ, HOSTSIM_DEVICES=str(n_dev))
    p = subprocess.run([hostsim] + golden.argv(case), capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    got = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
    assert got == golden.sam(case)
    used = [int(l.split()[4]) for l in p.stderr.split('\n') if l.startswith('[D::hostsim] device')]
    assert len(used) == n_dev and all(u >= 1 for u in used) and max(used) - min(used) <= 1
    bs = {}
    for l in p.stderr.split('\n'):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            bs[k] = bs.get(k, 0) + int(v)
    assert bs == golden.cases[case]['bsstat']
