#!/usr/bin/env python3
"""Golden fixtures for BASELINE configs C1 and C3 on the REFERENCE's own test genome (run once, in the build
container; needs /root/reference and oracle/_ref from `make -C oracle ref`).

  c1/BSB_test.fa.gz           the reference-held fixture tests/TestData/BSB_test.fa (6 contigs, 1.96 Mb; chr15 repeats
                              the first 5 kb of chr10), byte for byte
  c1/test_wgbs_masking.bed    tests/TestData/test_wgbs_masking.bed (the `-MR` fixture of tests/test_masked_alignment.py)
  c1/se100_1.fq.gz            reads of the REFERENCE simulator: `bsbolt Simulate -G BSB_test.fa -RL 100 -RD 5` (single end,
                              directional -- the PR1 set of BASELINE configs[0])
  c1c3.json                   md5 of every file the REFERENCE `bsbolt Index` writes for that genome in three modes -- whole
                              genome, `-MR` bed-masked, `-rrbs` (MspI, 30..500) -- i.e. BSB_ref.fa, the six `bwa index` files and
                              mappable_regions.bed (uncompressed); plus md5 / record count / BSStat of the SAM the REFERENCE
                              aligner prints for the C1 reads (argv of bsbolt/Utils/Launcher.py:75-115, -K fixed) and for the C3
                              reads (seeded RRBS simulator of bsbolt_b200/simulate.py, SE50, ~120 reads per start site) on the
                              RRBS database
"""
import gzip, hashlib, json, os, shutil, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import LAUNCHER_ARGS, PYSAM_STUB, RUN_REF, REF, gz, md5  # noqa: E402

WORK = os.path.join(os.environ.get('BSB_WORK', '/tmp/bsb_work'), 'c1c3')
TESTDATA = '/root/reference/tests/TestData'
C3_READS = dict(n_reads=120000, read_len=50, seed=33)


def ref_sam(db, fqs, extra):
    cmd = [REF + '/bwa', 'mem'] + LAUNCHER_ARGS + extra + [db] + fqs
    p = subprocess.run(cmd, check=True, capture_output=True, text=True)
    sam = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
    stats = {}
    for l in p.stderr.split('\n'):
        if l.startswith('BSStat '):
            k, v = l[7:].split(': ')
            stats[k] = stats.get(k, 0) + int(v)
    return dict(extra=extra, bsstat=stats, n_records=sum(1 for l in sam.split('\n') if l and l[0] != '@'),
                sam_md5=hashlib.md5(sam.encode()).hexdigest())


def main():
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(WORK + '/stub/pysam')
    open(WORK + '/stub/pysam/__init__.py', 'w').write(PYSAM_STUB)
    open(WORK + '/run_ref.py', 'w').write(RUN_REF % dict(stub=WORK + '/stub', ref=REF))
    run_ref = [sys.executable, WORK + '/run_ref.py']
    fa = WORK + '/BSB_test.fa'
    shutil.copy(TESTDATA + '/BSB_test.fa', fa)
    os.chmod(fa, 0o644)
    out = HERE + '/c1'
    os.makedirs(out, exist_ok=True)
    gz(fa, out + '/BSB_test.fa.gz')
    shutil.copy(TESTDATA + '/test_wgbs_masking.bed', out + '/test_wgbs_masking.bed')
    os.chmod(out + '/test_wgbs_masking.bed', 0o644)
    man = {'launcher_args': LAUNCHER_ARGS, 'fasta_md5': md5(fa), 'db': {}}
    modes = {'wgbs': [], 'masked': ['-MR', TESTDATA + '/test_wgbs_masking.bed'], 'rrbs': ['-rrbs']}
    for mode, extra in modes.items():
        db = f'{WORK}/db_{mode}'
        subprocess.run(run_ref + ['Index', '-G', fa, '-DB', db] + extra, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        man['db'][mode] = {f: md5(f'{db}/{f}') for f in ['BSB_ref.fa'] + [f'BSB_ref.fa.{e}' for e in ('amb', 'ann', 'bwt', 'sa', 'pac', 'opac')]}
        if mode == 'rrbs':
            man['db'][mode]['mappable_regions.bed'] = hashlib.md5(gzip.open(f'{db}/mappable_regions.bed.gz', 'rb').read()).hexdigest()
    # C1: the reference simulator's single-end 100 bp reads at depth 5
    subprocess.run(run_ref + ['Simulate', '-G', fa, '-O', f'{WORK}/se100', '-RL', '100', '-RD', '5', '-RS', '7'], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gz(f'{WORK}/se100_1.fq', out + '/se100_1.fq.gz')
    man['c1'] = ref_sam(f'{WORK}/db_wgbs/BSB_ref.fa', [f'{WORK}/se100_1.fq'], ['-K', '2000000'])
    man['c1']['fq_md5'] = md5(f'{WORK}/se100_1.fq')
    man['c1_undirectional'] = ref_sam(f'{WORK}/db_wgbs/BSB_ref.fa', [f'{WORK}/se100_1.fq'], ['-z', '-K', '2000000'])
    # C3: RRBS reads (seeded, generated again at test time) on the reference-built RRBS database
    from bsbolt_b200 import simulate
    names, seqs = simulate.read_fasta(fa)
    simulate.simulate_rrbs_reads(names, seqs, f'{WORK}/rrbs50.fq', **C3_READS)
    man['c3'] = ref_sam(f'{WORK}/db_rrbs/BSB_ref.fa', [f'{WORK}/rrbs50.fq'], ['-K', '1000000'])
    man['c3']['fq_md5'] = md5(f'{WORK}/rrbs50.fq')
    man['c3']['reads'] = C3_READS
    man['c3_on_masked_wgbs'] = ref_sam(f'{WORK}/db_masked/BSB_ref.fa', [f'{WORK}/rrbs50.fq'], ['-K', '1000000'])
    json.dump(man, open(HERE + '/c1c3.json', 'w'), indent=1, sort_keys=True)
    for k in ('c1', 'c1_undirectional', 'c3', 'c3_on_masked_wgbs'):
        print(k, man[k]['n_records'], man[k]['bsstat'])


if __name__ == '__main__':
    main()
