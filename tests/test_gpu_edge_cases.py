"""Edge cases of the alignment path, differential against a live run of the compiled reference (oracle/_ref/bwa travels to
the GPU box): empty and ragged inputs, reads shorter than a seed, ambiguous and lower-case bases, FASTA and gzip input,
and the options `bsbolt Align` can put on the `bwa mem` command line besides the defaults (-M -S -C -R -H -I -A ...).
Bit-exact like the other parity tests: header (all but @PG) and every record."""
import gzip
import os
import random
import subprocess

import pytest

from conftest import ROOT, first_diff, strip_pg

pytestmark = pytest.mark.gpu
BWA = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
COMP = str.maketrans('ACGTNacgtn', 'TGCANtgcan')


@pytest.fixture(scope='module')
def index(built, golden):
    from bsbolt_b200 import _native
    if not os.path.exists(BWA):
        pytest.skip('oracle/_ref/bwa did not travel to this box')
    ix = _native.Index(golden.idxbase, 0)
    yield ix
    ix.close()


def genome(golden):
    seqs, name = {}, None
    for line in open(os.path.join(golden.dir, 'genome.fa')):
        if line.startswith('>'):
            name = line[1:].split()[0]; seqs[name] = []
        else:
            seqs[name].append(line.strip())
    return {k: ''.join(v) for k, v in seqs.items()}


def bisulfite(s, rnd, rate=0.9):
    return ''.join('T' if c == 'C' and rnd.random() < rate else c for c in s)


def ragged_reads(golden, seed, paired):
    """reads of awkward lengths and contents; mates 350 bp apart where paired"""
    rnd = random.Random(seed)
    g = genome(golden)
    names = sorted(g)
    r1, r2 = [], []
    lens = [1, 2, 5, 18, 19, 20, 21, 30, 31, 32, 33, 47, 48, 49, 64, 65, 99, 100, 149, 150, 151, 200, 250, 16, 17, 255, 256]
    for k, L in enumerate(lens * 3):
        c = names[k % len(names)]
        L2 = lens[(k * 7 + 3) % len(lens)]
        p = rnd.randrange(0, len(g[c]) - 700)
        a = g[c][p:p + L]
        b = g[c][p + 350:p + 350 + L2][::-1].translate(COMP) if True else ''
        b = g[c][max(0, p + 350 - L2):p + 350][::-1].translate(COMP)
        if k % 4 == 1:                      # crick-strand pair
            a, b = b, a
        a, b = bisulfite(a, rnd), b.replace('G', 'A') if k % 3 else b
        kind = k % 11
        if kind == 3: a = 'N' * len(a)
        if kind == 4 and len(a) > 20: a = a[:len(a) // 2] + 'NNNN' + a[len(a) // 2 + 4:]
        if kind == 5: a = a.lower()
        if kind == 6: a = 'A' * len(a)
        if kind == 7 and len(a) > 10: a = a[:5] + a[5:10].lower() + a[10:]
        if kind == 8 and len(a) > 40:      # a few substitutions and a 2-base deletion
            a = list(a)
            for _ in range(3): a[rnd.randrange(len(a))] = rnd.choice('ACGT')
            a = ''.join(a); a = a[:len(a) // 3] + a[len(a) // 3 + 2:]
        q = lambda s: ''.join(chr(33 + rnd.randrange(2, 41)) for _ in s)
        r1.append((f'e{k}', a, q(a)))
        r2.append((f'e{k}', b, q(b)))
    return (r1, r2) if paired else (r1, None)


def write_fq(path, recs, fasta=False, comment=None, gz=False, suffix=''):
    op = gzip.open if gz else open
    with op(path, 'wt') as f:
        for n, s, q in recs:
            head = n + suffix + (' ' + comment if comment else '')
            f.write(f'>{head}\n{s}\n' if fasta else f'@{head}\n{s}\n+\n{q}\n')
    return str(path)


def both(index, golden, extra, fqs, tmp_path, tag):
    from bsbolt_b200 import _native
    argv = ['mem'] + golden.manifest['launcher_args'] + extra + [golden.idxbase] + fqs
    ref = subprocess.run([BWA] + argv, capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr[-1500:]
    out, log = tmp_path / f'{tag}.sam', tmp_path / f'{tag}.log'
    with open(out, 'w') as fo, open(log, 'w') as fl:
        rc, _ = _native.mem_main(argv, index=index, out_fd=fo.fileno(), log_fd=fl.fileno())
    assert rc == 0, _native.last_error()
    a, b = strip_pg(ref.stdout), strip_pg(open(out).read())
    assert a == b, f'{tag} {extra}: ' + first_diff(a, b)
    bs = lambda text: sorted(l for l in text.split('\n') if l.startswith('BSStat '))
    assert bs(ref.stderr) == bs(open(log).read()), f'{tag}: BSStat lines differ'
    return a


def test_ragged_single_end(index, golden, tmp_path):
    r1, _ = ragged_reads(golden, 5, False)
    fq = write_fq(tmp_path / 'r.fq', r1)
    for k, extra in enumerate((['-K', '3000'], ['-z', '-K', '100000'], ['-z', '-e', '0', '-K', '2500'], ['-L', '2,2', '-T', '0', '-K', '100000'])):
        sam = both(index, golden, extra, [fq], tmp_path, f'se{k}')
    assert sum(1 for l in sam.split('\n') if l and not l.startswith('@')) >= len(r1)


def test_ragged_paired_end(index, golden, tmp_path):
    r1, r2 = ragged_reads(golden, 6, True)
    fqs = [write_fq(tmp_path / 'r1.fq', r1, suffix='/1'), write_fq(tmp_path / 'r2.fq', r2, suffix='/2')]
    for k, extra in enumerate((['-K', '4000'], ['-z', '-K', '100000'], ['-S', '-K', '100000'], ['-I', '350,40', '-K', '5000'], ['-U', '3', '-K', '100000'])):
        both(index, golden, extra, fqs, tmp_path, f'pe{k}')


def test_empty_and_tiny_inputs(index, golden, tmp_path):
    empty = tmp_path / 'empty.fq'
    empty.write_text('')
    both(index, golden, ['-K', '1000'], [str(empty)], tmp_path, 'empty_se')
    both(index, golden, ['-K', '1000'], [str(empty), str(empty)], tmp_path, 'empty_pe')
    g = genome(golden)
    one = [('only', g['chr1'][1000:1100], 'I' * 100)]
    both(index, golden, [], [write_fq(tmp_path / 'one.fq', one)], tmp_path, 'one')
    mate = [('only', g['chr1'][1300:1400][::-1].translate(COMP), 'I' * 100)]
    both(index, golden, [], [write_fq(tmp_path / 'o1.fq', one), write_fq(tmp_path / 'o2.fq', mate)], tmp_path, 'one_pair')


def test_zero_length_reads(index, golden, tmp_path):
    g = genome(golden)
    recs = [('z0', '', ''), ('ok', g['chr2'][500:600], 'I' * 100), ('z1', '', ''), ('z2', '', '')]
    mates = [('z0', g['chr2'][800:900][::-1].translate(COMP), 'I' * 100), ('ok', '', ''), ('z1', '', ''), ('z2', 'ACGT', 'IIII')]
    fq1, fq2 = write_fq(tmp_path / 'z1.fq', recs), write_fq(tmp_path / 'z2.fq', mates)
    both(index, golden, [], [fq1], tmp_path, 'zero_se')
    both(index, golden, ['-z'], [fq1], tmp_path, 'zero_se_un')
    both(index, golden, [], [fq1, fq2], tmp_path, 'zero_pe')


def test_reads_beyond_the_supported_length_fail_loudly(index, golden, tmp_path):
    """reads longer than the per-warp shared-memory tiles (1200 bp): refuse, never differ silently"""
    from bsbolt_b200 import _native
    g = genome(golden)
    fq = write_fq(tmp_path / 'long.fq', [('long', g['chr1'][100:1400], 'I' * 1300)])
    argv = ['mem'] + golden.manifest['launcher_args'] + [golden.idxbase, fq]
    with open(tmp_path / 'o.sam', 'w') as fo, open(tmp_path / 'o.log', 'w') as fl:
        rc, _ = _native.mem_main(argv, index=index, out_fd=fo.fileno(), log_fd=fl.fileno())
    assert rc != 0 and '1200' in _native.last_error()


def long_reads(golden, seed, n=120):
    """pairs of 650-1200 bp with substitutions and indels: short seeds around the errors, so mem_seed_sw has work"""
    rnd = random.Random(seed)
    g = genome(golden)
    names = sorted(k for k in g if len(g[k]) > 6000)

    def mut(s, sub, indel):
        out = []
        for c in s:
            r = rnd.random()
            if r < sub: out.append(rnd.choice('ACGT'))
            elif r < sub + indel: continue
            elif r < sub + 2 * indel: out.append(c + rnd.choice('ACGT'))
            else: out.append(c)
        return ''.join(out)[:1200]
    r1, r2 = [], []
    for k in range(n):
        c = names[k % len(names)]
        L = rnd.choice([650, 700, 719, 720, 721, 760, 850, 1000, 1190])
        p = rnd.randrange(0, len(g[c]) - 2 * L - 400)
        a, b = g[c][p:p + L].upper(), g[c][p + L + 150:p + 2 * L + 150].upper()[::-1].translate(COMP)
        sub, ind = rnd.choice([0.0, 0.02, 0.05, 0.09]), rnd.choice([0, 0.003, 0.01])
        a, b = mut(bisulfite(a, rnd), sub, ind), mut(b.replace('G', 'A'), sub, ind)
        if k % 13 == 0:
            a = a[:len(a) // 2] + ''.join(rnd.choice('ACGT') for _ in range(len(a) // 2))
        r1.append((f'l{k}', a, 'I' * len(a)))
        r2.append((f'l{k}', b, 'F' * len(b)))
    return r1, r2


def test_long_reads_chained_seed_filter(index, golden, tmp_path):
    """mem_flt_chained_seeds (bwamem.c:602-619) on reads around and above its ~720 bp threshold, single and paired, and switched
    on for short reads by -W (min_l = 1.1 W instead of 5.5 ln l)"""
    r1, r2 = long_reads(golden, 10)
    fqs = [write_fq(tmp_path / 'l1.fq', r1), write_fq(tmp_path / 'l2.fq', r2)]
    both(index, golden, ['-K', '30000'], fqs[:1], tmp_path, 'long_se')
    both(index, golden, ['-z', '-K', '100000'], fqs[:1], tmp_path, 'long_se_un')
    both(index, golden, ['-K', '50000'], fqs, tmp_path, 'long_pe')
    both(index, golden, ['-W', '25', '-K', '50000'], fqs, tmp_path, 'long_pe_w25')
    s1, s2 = ragged_reads(golden, 11, True)
    sq = [write_fq(tmp_path / 's1.fq', s1), write_fq(tmp_path / 's2.fq', s2)]
    both(index, golden, ['-W', '1', '-k', '12', '-K', '50000'], sq, tmp_path, 'short_w1')
    both(index, golden, ['-W', '4', '-K', '50000'], sq, tmp_path, 'short_w4')


def test_input_formats(index, golden, tmp_path):
    r1, r2 = ragged_reads(golden, 7, True)
    base = both(index, golden, ['-K', '100000'], [write_fq(tmp_path / 'a1.fq', r1), write_fq(tmp_path / 'a2.fq', r2)], tmp_path, 'plain')
    gz = both(index, golden, ['-K', '100000'], [write_fq(tmp_path / 'g1.fq.gz', r1, gz=True), write_fq(tmp_path / 'g2.fq.gz', r2, gz=True)], tmp_path, 'gz')
    assert gz == base
    both(index, golden, ['-K', '100000'], [write_fq(tmp_path / 'f1.fa', r1, fasta=True), write_fq(tmp_path / 'f2.fa', r2, fasta=True)], tmp_path, 'fasta')


def test_launcher_options(index, golden, tmp_path):
    """every option bsbolt/Utils/Launcher.py:77-98 can add to the command line, away from its default"""
    r1, r2 = ragged_reads(golden, 8, True)
    fqs = [write_fq(tmp_path / 'c1.fq', r1, comment='BC:Z:ACGT'), write_fq(tmp_path / 'c2.fq', r2, comment='BC:Z:ACGT')]
    hdr = tmp_path / 'hdr.txt'
    hdr.write_text('@CO\tfirst extra line\n@CO\tsecond extra line\n')
    for k, extra in enumerate((['-M'], ['-j'], ['-C'], ['-R', r'@RG\tID:grp1\tSM:sample'], ['-H', '@CO\textra header line'], ['-H', str(hdr)],
                               ['-A', '2'], ['-A', '2', '-B', '6', '-O', '7,5', '-E', '2,1', '-L', '10,12'], ['-T', '40', '-h', '3,50'],
                               ['-k', '12', '-r', '1.0', '-c', '30', '-D', '0.3', '-m', '10', '-W', '5', '-y', '5'],
                               ['-d', '30', '-w', '20', '-Z', '0.5', '-z', '-l', '0.3', '-n', '2'])):
        both(index, golden, extra + ['-K', '50000'], fqs, tmp_path, f'opt{k}')


def test_smart_pairing(index, golden, tmp_path):
    """`-p`: interleaved ragged pairs with singletons and orphans mixed in, alone and with the options that change what the
    two per-batch calls see (-I statistics for the paired call only, comments, read group, a second file that is ignored)"""
    r1, r2 = ragged_reads(golden, 9, True)
    inter = []
    for k, (a, b) in enumerate(zip(r1, r2)):
        if k % 5 == 2:
            inter.append((f's{k}', a[1][::-1].translate(COMP), a[2]))   # a singleton between two pairs
        inter.append(a)
        if k % 6 != 4:
            inter.append(b)                                             # else: orphaned first mate
    fq = write_fq(tmp_path / 'inter.fq', inter, comment='BC:Z:ACGT')
    for k, extra in enumerate((['-p', '-K', '3000'], ['-p', '-K', '100000'], ['-p', '-z', '-K', '2000'], ['-p', '-I', '350,40', '-K', '5000'],
                               ['-p', '-C', '-R', r'@RG\tID:g\tSM:s', '-K', '4000'], ['-p', '-S', '-P', '-K', '100000'])):
        both(index, golden, extra, [fq], tmp_path, f'smart{k}')
    both(index, golden, ['-p', '-K', '100000'], [fq, fq], tmp_path, 'smart_two_files')
    single = write_fq(tmp_path / 'single.fq', inter[:1])
    both(index, golden, ['-p'], [single], tmp_path, 'smart_one_read')


def test_bwa_mem_options_beyond_the_launcher(index, golden, tmp_path):
    """`bsb_mem_main` takes a full `bwa mem` command line: the options `bsbolt Align` never sets (fastmap.c:113-190) behave like
    the reference's too -- -5 (mem_reorder_primary5), -a, -q, -u, -V, -P, -s/-G/-N/-X/-Q and the -x presets"""
    r1, r2 = ragged_reads(golden, 12, True)
    fqs = [write_fq(tmp_path / 'b1.fq', r1), write_fq(tmp_path / 'b2.fq', r2)]
    pe = [os.path.join(golden.dir, 'pe150uc_1.fq'), os.path.join(golden.dir, 'pe150uc_2.fq')]
    se = [os.path.join(golden.dir, 'se100c.fq')]
    for k, extra in enumerate((['-5'], ['-5', '-z'], ['-a'], ['-q'], ['-u'], ['-V'], ['-P'], ['-S', '-P'], ['-s', '3', '-G', '50', '-N', '2'],
                               ['-X', '0.2'], ['-Q', '20'], ['-Q', '0'])):
        both(index, golden, extra + ['-K', '50000'], fqs, tmp_path, f'bwa{k}')
    both(index, golden, ['-5', '-K', '150000'], pe, tmp_path, 'bwa_5_pe')
    both(index, golden, ['-5', '-M', '-K', '100000'], se, tmp_path, 'bwa_5_se')
    both(index, golden, ['-a', '-M', '-u', '-K', '100000'], se, tmp_path, 'bwa_amu_se')
    both(index, golden, ['-x', 'pacbio', '-K', '100000'], se, tmp_path, 'bwa_pacbio')
    long_fq = [os.path.join(golden.dir, 'long_se.fq')]
    both(index, golden, ['-x', 'ont2d', '-K', '100000'], long_fq, tmp_path, 'bwa_ont2d')
    both(index, golden, ['-x', 'intractg', '-5', '-K', '100000'], long_fq, tmp_path, 'bwa_intractg')


def test_differential_fuzz_on_the_device(index, golden, tmp_path):
    """tools/fuzz_hostsim.py --gpu: adversarial reads (chimeras, repeats, mis-oriented and overlapping mates, homopolymers, N runs)
    under random option sets, the product binary (bsbolt_b200/bwa over the C ABI) against the compiled reference, record by
    record and BSStat line by line. 70 runs of the same generator are logged in profiles/r01_fuzz_gpu_70_runs.log."""
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_hostsim.py'), '--gpu', '--runs', '8', '--reads', '400', '--seed', '5',
                        '--work', str(tmp_path / 'fz')], capture_output=True, text=True, errors='backslashreplace')
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-1000:]
    assert p.stdout.count('-> identical') == 8
