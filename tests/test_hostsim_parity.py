"""CPU-side unit tests of the kernel bodies (bsbolt_b200/csrc/bsb_*.h) through tests/hostsim, against
the committed golden SAM of the reference aligner. The same bodies are what the sm_100a kernels run; the
GPU tests (test_gpu_parity.py) repeat these comparisons through the C ABI on the device."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, first_diff, strip_pg
from smart_inter import SMART_CASES, write_interleaved

HOSTSIM = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim')
CASES = ['se100', 'se100_un', 'se50_clip', 'pe150', 'pe150_un', 'pe150_un_sp0', 'pe150_opts']


@pytest.mark.parametrize('seeding', ['nested', 'two_item'])
@pytest.mark.parametrize('case', CASES)
def test_kernel_bodies_match_reference_sam(built, golden, case, seeding):
    """`two_item` = the seeding form the product kernel runs (bsb_seed3.h); `nested` = the direct restatement."""
    env = dict(os.environ)
    if seeding == 'two_item':
        env['BSB_HOSTSIM_SEED_V3'] = '1'
    p = subprocess.run([HOSTSIM] + golden.argv(case), capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    mine, want = strip_pg(p.stdout), golden.sam(case)
    assert mine == want, first_diff(want, mine)
    stats = {}
    for l in p.stderr.split('\n'):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            stats[k] = stats.get(k, 0) + int(v)
    assert stats == golden.cases[case]['bsstat']


@pytest.mark.parametrize('case', sorted(SMART_CASES))
def test_smart_pairing_matches_reference_sam(built, golden, tmp_path, case):
    """`-p` (fastmap.c:38-57, bseq_classify bwa.c:147-165): one interleaved file, single-end entries and name-matched
    pairs aligned as two calls per batch, the arbiter fed with untouched statistics -- against the reference's SAM."""
    import gzip
    import json
    man = json.load(open(os.path.join(GOLDEN, 'smart_golden.json')))[case]
    fq = write_interleaved(golden.dir, tmp_path / 'smart_inter.fq')
    argv = ['mem'] + golden.manifest['launcher_args'] + man['extra'] + [golden.idxbase, fq]
    p = subprocess.run([HOSTSIM] + argv, capture_output=True, text=True, env=dict(os.environ, BSB_HOSTSIM_SEED_V3='1'))
    assert p.returncode == 0, p.stderr[-2000:]
    mine, want = strip_pg(p.stdout), gzip.open(os.path.join(GOLDEN, case + '.sam.gz'), 'rt').read()
    assert mine == want, first_diff(want, mine)
    stats = {}
    for l in p.stderr.split('\n'):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            stats[k] = stats.get(k, 0) + int(v)
    assert stats == man['bsstat']


@pytest.mark.parametrize('case', ['long_se', 'long_se_w30'])
def test_chained_seed_filter_matches_reference_sam(built, golden, case):
    """mem_flt_chained_seeds / mem_seed_sw (bwamem.c:575-619): reads of 690-1200 bp (the filter starts at about 720 bp), and
    the same reads with -W 30, which switches it on for every length"""
    import gzip
    import json
    man = json.load(open(os.path.join(GOLDEN, 'long_golden.json')))[case]
    argv = ['mem'] + golden.manifest['launcher_args'] + man['extra'] + [golden.idxbase] + [os.path.join(golden.dir, f) for f in man['fq']]
    p = subprocess.run([HOSTSIM] + argv, capture_output=True, text=True, env=dict(os.environ, BSB_HOSTSIM_SEED_V3='1'))
    assert p.returncode == 0, p.stderr[-2000:]
    mine, want = strip_pg(p.stdout), gzip.open(os.path.join(GOLDEN, case + '.sam.gz'), 'rt').read()
    assert mine == want, first_diff(want, mine)


def test_reference_binary_reproduces_golden(built, golden):
    """Pins oracle/_ref (the compiled reference) to the committed vectors."""
    bwa = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
    if not os.path.exists(bwa):
        pytest.skip('oracle/_ref/bwa not built (no /root/reference and no prebuilt copy)')
    for case in ('se100', 'pe150_un'):
        p = subprocess.run([bwa] + golden.argv(case), capture_output=True, text=True)
        assert p.returncode == 0
        assert strip_pg(p.stdout) == golden.sam(case)


def test_differential_fuzz_of_the_kernel_bodies(built, tmp_path):
    """tools/fuzz_hostsim.py: adversarial reads under random option sets, kernel bodies on the CPU against the compiled
    reference (skipped where oracle/_ref/bwa is absent); with --bam every run is also written as BAM through the emulated device
    stage (arbiter, record encoder, deflate) and through the host encoder, and the inflated streams must be identical"""
    import sys
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'bwa')):
        pytest.skip('oracle/_ref/bwa not built')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_hostsim.py'), '--runs', '5', '--reads', '300', '--seed', '3', '--bam',
                        '--work', str(tmp_path / 'fz')], capture_output=True, text=True, errors='backslashreplace')
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-1000:]
    assert p.stdout.count('-> identical') == 5 and p.stdout.count('BAM stream identical') == 5
