#!/usr/bin/env python3
"""Golden SAM of the REFERENCE aligner (oracle/_ref/bwa, built from /root/reference) for smart pairing (`bwa mem -p`,
`bsbolt Align -p`) on the interleaved file tests/smart_inter.py derives from the committed FASTQ fixtures.
Writes smart_p.sam.gz, smart_p_un.sam.gz and smart_golden.json. Run once in the build container."""
import gzip, hashlib, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from smart_inter import SMART_CASES, write_interleaved  # noqa: E402


def main():
    man = json.load(open(os.path.join(HERE, 'golden.json')))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(d + '/db')
        for f in os.listdir(HERE + '/db'):
            with gzip.open(f'{HERE}/db/{f}', 'rb') as i, open(f'{d}/db/{f[:-3]}', 'wb') as o:
                shutil.copyfileobj(i, o)
        for f in ('pe150c_1.fq', 'pe150c_2.fq', 'se100c.fq'):
            with gzip.open(f'{HERE}/{f}.gz', 'rb') as i, open(f'{d}/{f}', 'wb') as o:
                shutil.copyfileobj(i, o)
        fq = write_interleaved(d, d + '/smart_inter.fq')
        for case, extra in SMART_CASES.items():
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
            out[case] = dict(extra=extra, bsstat=stats, n_records=sum(1 for l in sam.split('\n') if l and l[0] != '@'),
                             sam_md5=hashlib.md5(sam.encode()).hexdigest(), fq_md5=hashlib.md5(open(fq, 'rb').read()).hexdigest())
            print(case, out[case])
    json.dump(out, open(HERE + '/smart_golden.json', 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
