#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ (run once, in the build container).

  genome.fa.gz        synthetic 6-contig reference (seeded numpy): i.i.d. bases with CpG-rich islands, one
                      duplicated 3-kb segment (exercises XA / YC:i:1 / MAPQ 0), one N-run, lower-case bases
  db/BSB_ref.fa.*     index written by the REFERENCE `bwa index -a bwtsw` (oracle/_ref/bwa, built from
                      /root/reference by oracle/Makefile), gzip-compressed
  *.fq.gz             reads simulated by the REFERENCE `bsbolt Simulate` (python package imported from
                      /root/reference with its wgsim built into oracle/_ref) + a seeded corruption pass
                      (substitutions, chimeric tails, N's) so that clipping and mate rescue fire
  *.sam.gz            what the REFERENCE aligner prints for them with the argv `bsbolt Align` builds
                      (bsbolt/Utils/Launcher.py:75-115) plus a fixed -K; @PG line removed
  golden.json         argv, md5 of every file, BSStat counters per case

Needs /root/reference and oracle/_ref (make -C oracle ref). Nothing here runs on the GPU box.
"""
import gzip, hashlib, json, os, random, shutil, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, 'oracle', '_ref')
WORK = os.path.join(os.environ.get('BSB_WORK', '/tmp/bsb_work'), 'golden')
LAUNCHER_ARGS = ('-Y -A 1 -B 4 -D 0.5 -E 1,1 -L 30,30 -T 10 -U 17 -W 0 -c 500 -d 100 -k 19 -m 50 -r 1.5 -t 1 -w 100 -y 20 '
                 '-O 6,6 -h 100,200 -e 0.1 -l 0.5 -n 5 -Z 0.95').split()

RUN_REF = r'''
import sys
sys.path.insert(0, %(stub)r); sys.path.insert(0, '/root/reference')
import bsbolt.Utils.UtilityFunctions as U
U.get_external_paths = lambda: (%(ref)r + '/bwa', %(ref)r + '/wgsim', %(ref)r + '/stream_bam')
from bsbolt.Utils.Parser import parser
from bsbolt.Utils import Launcher
Launcher.bwa_path, Launcher.wgsim_path, Launcher.stream_bam = U.get_external_paths()
args = parser.parse_args(sys.argv[1:])
Launcher.bsb_launch[args.subparser_name](args)
'''

PYSAM_STUB = '''
import sys
class _Any:
    def __getattr__(self, k): return _Any()
    def __call__(self, *a, **k): return _Any()
class _Mod(type(sys)):
    def __getattr__(self, k): return _Any()
sys.modules[__name__].__class__ = _Mod
'''


def make_genome(path):
    rng = np.random.default_rng(20240517)
    lens = [90000, 70000, 60000, 40000, 30000, 3000]
    seqs = []
    for i, n in enumerate(lens[:-1]):
        s = rng.choice(np.frombuffer(b'ACGT', dtype='S1'), size=n, p=[.29, .21, .21, .29]).astype('S1')
        for _ in range(n // 6000):  # CpG islands
            p = int(rng.integers(0, n - 400))
            isl = rng.choice(np.frombuffer(b'ACGT', dtype='S1'), size=400, p=[.15, .35, .35, .15])
            s[p:p + 400] = isl
        seqs.append(s)
    seqs[1][20000:20060] = b'N'                      # N-run (random bases in pac/opac)
    low = seqs[2][5000:5200]
    seqs[2][5000:5200] = np.char.lower(low)          # soft-masked stretch
    seqs.append(seqs[0][10000:13000].copy())         # chr6 duplicates chr1:10000-13000
    with open(path, 'w') as f:
        for i, s in enumerate(seqs):
            f.write(f'>chr{i + 1}\n')
            t = b''.join(s.tolist()).decode()
            for k in range(0, len(t), 60):
                f.write(t[k:k + 60] + '\n')


def corrupt(inp, out, frac, nfrac, seed):
    rnd = random.Random(seed)
    with open(inp) as f, open(out, 'w') as o:
        while True:
            h = f.readline()
            if not h: break
            s = list(f.readline().rstrip('\n')); p = f.readline(); q = f.readline()
            r = rnd.random()
            if r < frac:
                for i in range(len(s)):
                    if rnd.random() < 0.15: s[i] = rnd.choice('ACGT')
            elif r < frac * 1.5:
                for i in range(len(s) // 2, len(s)): s[i] = rnd.choice('ACGT')
            elif r < frac * 2:
                for i in range(0, len(s) // 3): s[i] = rnd.choice('ACGT')
            if rnd.random() < nfrac:
                for _ in range(rnd.randint(1, 4)): s[rnd.randrange(len(s))] = 'N'
            o.write(h + ''.join(s) + '\n' + p + q)


def md5(path):
    return hashlib.md5(open(path, 'rb').read()).hexdigest()


def gz(src, dst):
    with open(src, 'rb') as f, gzip.GzipFile(dst, 'wb', mtime=0) as g:
        shutil.copyfileobj(f, g)


def main():
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(WORK + '/stub/pysam')
    open(WORK + '/stub/pysam/__init__.py', 'w').write(PYSAM_STUB)
    open(WORK + '/run_ref.py', 'w').write(RUN_REF % dict(stub=WORK + '/stub', ref=REF))
    run_ref = [sys.executable, WORK + '/run_ref.py']
    fa = WORK + '/genome.fa'
    make_genome(fa)
    subprocess.run(run_ref + ['Index', '-G', fa, '-DB', WORK + '/db'], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sims = {'se100': ['-RL', '100', '-RD', '1', '-RS', '7'],
            'pe150': ['-PE', '-RL', '150', '-RD', '2', '-RS', '5'],
            'pe150u': ['-PE', '-RL', '150', '-RD', '2', '-RS', '11', '-U'],
            'se50': ['-RL', '50', '-RD', '1', '-RS', '3']}
    for name, a in sims.items():
        subprocess.run(run_ref + ['Simulate', '-G', fa, '-O', f'{WORK}/{name}'] + a, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    corrupt(f'{WORK}/pe150_1.fq', f'{WORK}/pe150c_1.fq', 0.04, 0.05, 1)
    corrupt(f'{WORK}/pe150_2.fq', f'{WORK}/pe150c_2.fq', 0.12, 0.05, 2)
    corrupt(f'{WORK}/pe150u_1.fq', f'{WORK}/pe150uc_1.fq', 0.04, 0.05, 3)
    corrupt(f'{WORK}/pe150u_2.fq', f'{WORK}/pe150uc_2.fq', 0.12, 0.05, 4)
    corrupt(f'{WORK}/se100_1.fq', f'{WORK}/se100c.fq', 0.08, 0.05, 5)
    cases = {
        'se100': dict(fq=['se100c.fq'], extra=['-K', '100000']),
        'se100_un': dict(fq=['se100c.fq'], extra=['-z', '-K', '100000']),
        'se50_clip': dict(fq=['se50_1.fq'], extra=['-L', '1,1', '-K', '50000']),
        'pe150': dict(fq=['pe150c_1.fq', 'pe150c_2.fq'], extra=['-K', '200000']),
        'pe150_un': dict(fq=['pe150uc_1.fq', 'pe150uc_2.fq'], extra=['-z', '-K', '200000']),
        'pe150_un_sp0': dict(fq=['pe150uc_1.fq', 'pe150uc_2.fq'], extra=['-z', '-e', '0', '-K', '150000']),
        'pe150_opts': dict(fq=['pe150c_1.fq', 'pe150c_2.fq'], extra=['-L', '5,5', '-k', '14', '-c', '60', '-K', '300000']),
    }
    manifest = {'launcher_args': LAUNCHER_ARGS, 'cases': {}, 'md5': {}}
    out = HERE
    shutil.rmtree(out + '/db', ignore_errors=True)
    os.makedirs(out + '/db')
    gz(fa, out + '/genome.fa.gz')
    for ext in ('amb', 'ann', 'bwt', 'sa', 'pac', 'opac'):
        gz(f'{WORK}/db/BSB_ref.fa.{ext}', f'{out}/db/BSB_ref.fa.{ext}.gz')
        manifest['md5'][f'db/BSB_ref.fa.{ext}'] = md5(f'{WORK}/db/BSB_ref.fa.{ext}')
    done = set()
    for name, c in cases.items():
        for fq in c['fq']:
            if fq not in done:
                gz(f'{WORK}/{fq}', f'{out}/{fq}.gz'); done.add(fq)
                manifest['md5'][fq] = md5(f'{WORK}/{fq}')
        argv = ['mem'] + LAUNCHER_ARGS + c['extra'] + ['BSB_ref.fa'] + c['fq']
        cmd = [REF + '/bwa', 'mem'] + LAUNCHER_ARGS + c['extra'] + [WORK + '/db/BSB_ref.fa'] + [f'{WORK}/{f}' for f in c['fq']]
        p = subprocess.run(cmd, check=True, capture_output=True, text=True)
        sam = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
        with gzip.GzipFile(f'{out}/{name}.sam.gz', 'wb', mtime=0) as g:
            g.write(sam.encode())
        stats = {}
        for l in p.stderr.split('\n'):
            if l.startswith('BSStat '):
                k, v = l[7:].split(': ')
                stats[k] = stats.get(k, 0) + int(v)
        manifest['cases'][name] = dict(argv=argv, fq=c['fq'], extra=c['extra'], bsstat=stats,
                                       n_records=sum(1 for l in sam.split('\n') if l and l[0] != '@'),
                                       sam_md5=hashlib.md5(sam.encode()).hexdigest())
        print(name, manifest['cases'][name]['n_records'], stats)
    json.dump(manifest, open(out + '/golden.json', 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
