#!/usr/bin/env python3
"""Golden vectors for reads long enough for mem_flt_chained_seeds / mem_seed_sw (bwamem.c:575-619; 5.5 ln(l) <= 0.05 l, about
720 bp and more) and for the same filter switched on by -W: long_se.fq.gz (seeded reads of 690-1200 bp cut from genome.fa.gz,
bisulfite-converted, with substitutions, indels and chimeric tails) and what the REFERENCE aligner (oracle/_ref/bwa, built from
/root/reference) prints for it. Writes long_se.fq.gz, long_se.sam.gz, long_se_w30.sam.gz and long_golden.json."""
import gzip, hashlib, json, os, random, shutil, subprocess, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
COMP = str.maketrans('ACGTN', 'TGCAN')
CASES = {'long_se': ['-K', '40000'], 'long_se_w30': ['-W', '30', '-K', '1000000']}


def make_reads(path):
    rnd = random.Random(42)
    g, name = {}, None
    for l in gzip.open(os.path.join(HERE, 'genome.fa.gz'), 'rt'):
        if l[0] == '>':
            name = l[1:].split()[0]; g[name] = []
        else:
            g[name].append(l.strip())
    g = {k: ''.join(v).upper() for k, v in g.items()}
    names = sorted(k for k in g if len(g[k]) > 5000)

    def mut(s, sub, indel):
        out = []
        for c in s:
            r = rnd.random()
            if r < sub: out.append(rnd.choice('ACGT'))
            elif r < sub + indel: continue
            elif r < sub + 2 * indel: out.append(c); out.append(rnd.choice('ACGT'))
            else: out.append(c)
        return ''.join(out)

    def bs(s):
        return ''.join('T' if c == 'C' and rnd.random() < 0.9 else c for c in s)
    with open(path, 'w') as f:
        for i in range(160):
            c = rnd.choice(names)
            L = rnd.choice([690, 700, 710, 719, 720, 725, 730, 750, 800, 900, 1000, 1100, 1200])
            p = rnd.randrange(0, len(g[c]) - L)
            s = g[c][p:p + L]
            sub = rnd.choice([0.0, 0.01, 0.03, 0.06, 0.10]); ind = rnd.choice([0, 0.002, 0.01])
            r = bs(s) if i % 2 == 0 else bs(s[::-1].translate(COMP))
            r = mut(r, sub, ind)[:1200]
            if i % 17 == 0:
                r = r[:len(r) // 2] + ''.join(rnd.choice('ACGT') for _ in range(len(r) // 2))
            f.write(f'@L{i}\n{r}\n+\n{"I" * len(r)}\n')


def main():
    man = json.load(open(os.path.join(HERE, 'golden.json')))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(d + '/db')
        for f in os.listdir(HERE + '/db'):
            with gzip.open(f'{HERE}/db/{f}', 'rb') as i, open(f'{d}/db/{f[:-3]}', 'wb') as o:
                shutil.copyfileobj(i, o)
        fq = d + '/long_se.fq'
        make_reads(fq)
        with open(fq, 'rb') as i, gzip.GzipFile(HERE + '/long_se.fq.gz', 'wb', mtime=0) as o:
            shutil.copyfileobj(i, o)
        for case, extra in CASES.items():
            p = subprocess.run([ROOT + '/oracle/_ref/bwa', 'mem'] + man['launcher_args'] + extra + [d + '/db/BSB_ref.fa', fq],
                               check=True, capture_output=True, text=True)
            sam = ''.join(l + '\n' for l in p.stdout.split('\n') if l and not l.startswith('@PG'))
            with gzip.GzipFile(f'{HERE}/{case}.sam.gz', 'wb', mtime=0) as g:
                g.write(sam.encode())
            stats = {}
            for l in p.stderr.split('\n'):
                if l.startswith('BSStat '):
                    k, v = l[7:].split(': ')
                    stats[k] = stats.get(k, 0) + int(v)
            out[case] = dict(extra=extra, fq=['long_se.fq'], bsstat=stats, n_records=sum(1 for l in sam.split('\n') if l and l[0] != '@'),
                             sam_md5=hashlib.md5(sam.encode()).hexdigest())
            print(case, out[case])
    json.dump(out, open(HERE + '/long_golden.json', 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
