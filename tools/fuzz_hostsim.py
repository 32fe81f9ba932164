#!/usr/bin/env python3
"""Development aid: differential fuzzing of the kernel bodies (tests/hostsim, CPU) or, with --gpu, of the product on the device against the compiled reference
(oracle/_ref/bwa) on adversarial reads cut from the golden genome -- chimeras, repeats, overlapping and mis-oriented mates,
heavy mutation, homopolymers, N runs -- under random option sets. Prints one line per run; exits 1 at the first difference.

python tools/fuzz_hostsim.py [--runs 20] [--seed 1] [--work /tmp/bsb_fuzz]"""
import argparse, gzip, json, os, random, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'tests', 'golden')
COMP = str.maketrans('ACGTNacgtn', 'TGCANtgcan')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--runs', type=int, default=20); ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--work', default='/tmp/bsb_fuzz'); ap.add_argument('--reads', type=int, default=600)
    ap.add_argument('--first', type=int, default=0, help='first run to execute (a run is determined by seed and run number)')
    ap.add_argument('--keep', action='store_true', help='keep ref.sam / mine.sam / logs of a differing run under --work')
    ap.add_argument('--long', action='store_true', help='mix in reads of 700-1200 bp (mem_flt_chained_seeds territory) and smart pairing (-p)')
    ap.add_argument('--bam', action='store_true', help='also write the run as BAM twice on the CPU harness -- through the emulated device stage (arbiter, bam_record, bgzf_block) and through the host encoder -- and compare the inflated streams')
    ap.add_argument('--gpu', action='store_true', help='run the product (bsbolt_b200/bwa, the bwa-compatible binary over the C ABI) on cuda:0 instead of the CPU harness')
    a = ap.parse_args()
    os.makedirs(a.work + '/db', exist_ok=True)
    for f in os.listdir(G + '/db'):
        with gzip.open(f'{G}/db/{f}', 'rb') as i, open(f'{a.work}/db/{f[:-3]}', 'wb') as o:
            shutil.copyfileobj(i, o)
    g, name = {}, None
    for l in gzip.open(G + '/genome.fa.gz', 'rt'):
        if l[0] == '>': name = l[1:].split()[0]; g[name] = []
        else: g[name].append(l.strip())
    g = {k: ''.join(v) for k, v in g.items()}
    names = sorted(g)
    base = json.load(open(G + '/golden.json'))['launcher_args']
    CAP = 1200 if a.long else 650
    for run in range(a.first, a.runs):
        rnd = random.Random(a.seed * 1000 + run)

        def locus(L):
            c = rnd.choice(names)
            if rnd.random() < 0.15: c, lo, hi = 'chr1', 9800, 13200   # the duplicated segment (chr6 = chr1:10000-13000)
            else: lo, hi = 0, len(g[c])
            L = min(L, hi - lo - 1)
            p = rnd.randrange(lo, hi - L)
            return c, p, g[c][p:p + L]

        def mutate(s):
            sub, ind = rnd.choice([0, 0, .01, .03, .08, .15]), rnd.choice([0, 0, .003, .01, .03])
            out = []
            for ch in s:
                r = rnd.random()
                if r < sub: out.append(rnd.choice('ACGTN' if rnd.random() < .1 else 'ACGT'))
                elif r < sub + ind: continue
                elif r < sub + 2 * ind: out.append(ch + rnd.choice('ACGT') * rnd.randint(1, 4))
                else: out.append(ch)
            return ''.join(out)

        def convert(s, pattern):
            rate = rnd.choice([1.0, .95, .7, .3])
            frm, to = ('C', 'T') if pattern == 0 else ('G', 'A')
            return ''.join(to if ch.upper() == frm and rnd.random() < rate else ch for ch in s)

        def one(L):
            kind = rnd.random()
            if kind < .06: return rnd.choice('ACGT') * L
            if kind < .10: return ''.join(rnd.choice('ACGT') for _ in range(L))
            c, p, s = locus(L)
            if kind < .25:                         # chimera of two loci
                s = s[:len(s) // 2] + locus(L - len(s) // 2)[2]
            if rnd.random() < .5: s = s[::-1].translate(COMP)
            s = mutate(convert(s, rnd.randrange(2)))
            if rnd.random() < .05 and len(s) > 30: s = s[:10] + 'N' * rnd.randint(1, 12) + s[20:]
            return s[:CAP] or 'A'
        paired = rnd.random() < .6
        r1, r2 = [], []
        for k in range(a.reads):
            L = rnd.choice([20, 25, 36, 50, 75, 100, 101, 125, 150, 150, 150, 200, 250, 300, 400])
            if a.long and rnd.random() < .3: L = rnd.choice([700, 719, 720, 760, 900, 1100, 1200])
            if paired:
                c, p, s = locus(L * 2 + rnd.choice([-L, 0, 50, 200, 400, 700]) if L * 3 < 60000 else L * 2)
                m1, m2 = s[:L], s[-L:][::-1].translate(COMP)
                o = rnd.random()
                if o < .1: m2 = m2[::-1].translate(COMP)           # same orientation
                elif o < .15: m1, m2 = m2, m1                       # outward facing
                elif o < .25: m2 = one(L)                           # unrelated mate
                if rnd.random() < .5: m1, m2 = mutate(convert(m1, 0)), mutate(convert(m2, 1))
                else: m1, m2 = mutate(convert(m2, 0)), mutate(convert(m1, 1))
                m1, m2 = m1[:CAP] or 'A', m2[:CAP] or 'A'
                r1.append((f'f{k}', m1)); r2.append((f'f{k}', m2))
            else:
                r1.append((f'f{k}', one(L)))
        fqs = []
        for tag, rs in (('1', r1), ('2', r2)):
            if rs:
                path = f'{a.work}/r{tag}.fq'
                with open(path, 'w') as f:
                    for n, s in rs:
                        if a.long and rnd.random() < .08 and len(s) > 20:   # a soft-masked stretch (conversion is upper-case only)
                            p0 = rnd.randrange(len(s) - 10); s = s[:p0] + s[p0:p0 + rnd.randint(1, 40)].lower() + s[p0 + 40:]
                        cm = ' BC:Z:' + ''.join(rnd.choice('ACGT') for _ in range(6)) if a.long and rnd.random() < .3 else ''
                        f.write(f'@{n}{"/" + tag if a.long and len(fqs) + 1 == int(tag) and rnd.random() < .5 else ""}{cm}\n{s}\n+\n{"".join(chr(33 + rnd.randrange(2, 41)) for _ in s)}\n')
                fqs.append(path)
        extra = ['-K', str(rnd.choice([3000, 20000, 100000, 10000000]))]
        if a.long and paired and rnd.random() < .3:   # smart pairing: one interleaved file, some mates dropped
            with open(f'{a.work}/inter.fq', 'w') as f:
                r1l, r2l = open(fqs[0]).read().split('\n'), open(fqs[1]).read().split('\n')
                for k in range(0, len(r1l) - 3, 4):
                    f.write('\n'.join(r1l[k:k + 4]) + '\n')
                    if rnd.random() > .1: f.write('\n'.join(r2l[k:k + 4]) + '\n')
            fqs = [f'{a.work}/inter.fq']
            extra.append('-p')
        for opt, vals in (('-z', [None]), ('-C', [None]), ('-M', [None]), ('-S', [None]), ('-P', [None]), ('-5', [None]), ('-a', [None]),
                          ('-k', ['10', '14', '25']), ('-c', ['5', '50']), ('-T', ['0', '30']), ('-L', ['0,0', '5,9']), ('-U', ['0', '40']),
                          ('-w', ['5', '30']), ('-d', ['20']), ('-r', ['0.8', '3']), ('-y', ['3', '0']), ('-A', ['2']), ('-B', ['2', '9']),
                          ('-O', ['3,9', '12,2']), ('-E', ['3,2']), ('-D', ['0.1', '0.9']), ('-W', ['1', '3', '8']), ('-m', ['2']),
                          ('-e', ['0', '0.3']), ('-h', ['2,5']), ('-Z', ['0.5']), ('-I', ['300,50', '150,20,400,10'])):
            if rnd.random() < .12:
                extra += [opt] if vals == [None] else [opt, rnd.choice(vals)]
        argv = ['mem'] + base + extra + [a.work + '/db/BSB_ref.fa'] + fqs
        ref = subprocess.run([ROOT + '/oracle/_ref/bwa'] + argv, capture_output=True, text=True, errors='backslashreplace')
        me = subprocess.run([ROOT + ('/bsbolt_b200/bwa' if a.gpu else '/tests/hostsim/hostsim')] + argv, capture_output=True, text=True, errors='backslashreplace', env=dict(os.environ, BSB_HOSTSIM_SEED_V3='1'))
        # records at the strand boundary (NM:i:4194303): the reference prints MD from memory it never wrote, i.e. arbitrary
        # bytes up to the first NUL, tabs included; this build prints an empty MD. Masked for the comparison.
        def strip(t):
            out = []
            for l in t.split('\n'):
                if not l or l.startswith('@PG'):
                    continue
                if '\tNM:i:4194303\t' in l:
                    f = l.split('\t')
                    k0 = next(i for i, x in enumerate(f) if x.startswith('NM:i:'))   # the undefined bytes may hold tabs: drop
                    k1 = next(i for i, x in enumerate(f) if i > k0 and x.startswith(('XC:i:', 'AS:i:')))   # everything between NM and XC/AS
                    l = '\t'.join(f[:k0 + 1] + f[k1:])
                out.append(l)
            return out
        x, y = strip(ref.stdout), strip(me.stdout)
        bs = lambda t: sorted(l for l in t.split('\n') if l.startswith('BSStat'))
        ok = ref.returncode == 0 and me.returncode == 0 and x == y and bs(ref.stderr) == bs(me.stderr)
        bam_note = ''
        if ok and a.bam and not a.gpu:
            raws, logs = [], []
            for host in (False, True):
                env = dict(os.environ, BSB_HOSTSIM_SEED_V3='1', HOSTSIM_BAM=f'{a.work}/fuzz_{int(host)}.bam')
                if host: env['HOSTSIM_BAM_HOST'] = '1'
                r = subprocess.run([ROOT + '/tests/hostsim/hostsim'] + argv, capture_output=True, text=True, errors='backslashreplace', env=env)
                if r.returncode: print('BAM run failed:', r.stderr[-400:]); sys.exit(1)
                raws.append(gzip.open(f'{a.work}/fuzz_{int(host)}.bam', 'rb').read()); logs.append(bs(r.stderr))
            if raws[0] != raws[1] or logs[0] != logs[1] or logs[0] != bs(me.stderr):
                print(f'run {run}: BAM streams differ (device stage {len(raws[0])} bytes, host encoder {len(raws[1])} bytes), BSStat equal {logs[0] == logs[1]}'); print('argv:', ' '.join(argv)); sys.exit(1)
            bam_note = f', BAM stream identical ({len(raws[0])} bytes)'
        print(f'run {run}: {"PE" if paired else "SE"} {len(x)} records {" ".join(extra)} -> {"identical" if ok else "DIFFERENT"}{bam_note}', flush=True)
        if not ok:
            print('ref rc', ref.returncode, 'mine rc', me.returncode, 'records', len(x), len(y), 'bsstat equal', bs(ref.stderr) == bs(me.stderr), me.stderr[-300:])
            if a.keep:
                for nm, t in (('ref.sam', ref.stdout), ('mine.sam', me.stdout), ('ref.log', ref.stderr), ('mine.log', me.stderr)):
                    open(f'{a.work}/{nm}', 'w').write(t)
                print('argv:', ' '.join(argv))
            for u, v in zip(x, y):
                if u != v:
                    print('ref :', u[:400]); print('mine:', v[:400]); break
            sys.exit(1)


if __name__ == '__main__':
    main()
