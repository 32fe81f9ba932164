#!/usr/bin/env python3
"""Known-answer vectors for the hot-path primitives, produced by calling the REAL reference functions
(oracle/_ref/libbwa_ref.so = the reference's own objects, built by oracle/Makefile) through ctypes on the
golden index: bwt_occ4, bwt_extend, bwt_smem1, bwt_seed_strategy1, bwt_sa, ksw_extend2, ksw_global2.
Output: tests/golden/primitives.json.gz (committed). Run in the build container only."""
import ctypes as C
import gzip, json, os, random, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from reflib import RefLib, unpack_index  # noqa: E402


def main():
    d = unpack_index()
    ref = RefLib(os.path.join(d, 'BSB_ref.fa'))
    rnd = random.Random(2024)
    out = {'occ4': [], 'extend': [], 'smem': [], 'seed_forward': [], 'sa': [], 'extend2': [], 'global2': []}
    n = ref.seq_len
    for _ in range(300):
        k = rnd.choice([rnd.randrange(0, n + 1), ref.primary, ref.primary - 1, ref.primary + 1, 0, n, 2 ** 64 - 1])
        out['occ4'].append([k, ref.occ4(k)])
    reads = []
    for l in gzip.open(os.path.join(HERE, 'se100c.fq.gz'), 'rt').read().split('\n')[1::4][:60]:
        reads.append(l)
    for r in reads:
        pat = rnd.randrange(2)
        q = [ref.code(c if not (c == 'CG'[pat]) else 'TA'[pat]) for c in r]
        x = rnd.randrange(len(q))
        mi = rnd.choice([1, 1, 2, 5])
        out['smem'].append(dict(q=q, x=x, min_intv=mi, res=ref.smem1(q, x, mi)))
        out['seed_forward'].append(dict(q=q, x=x, res=ref.seed_strategy1(q, x, 19, 20)))
        ik = ref.set_intv(q[x])
        for b in (0, 1):
            out['extend'].append(dict(ik=ik, is_back=b, ok=ref.extend(ik, b)))
    for _ in range(300):
        k = rnd.randrange(1, n + 1)
        out['sa'].append([k, ref.sa(k)])
    mat = ref.scmat(1, 4)
    for _ in range(150):
        ql, tl = rnd.randrange(1, 120), rnd.randrange(1, 160)
        t = [rnd.randrange(4) for _ in range(tl)]
        q = [(t[i] if i < tl and rnd.random() > 0.08 else rnd.randrange(5)) for i in range(ql)]
        if rnd.random() < 0.3:
            p = rnd.randrange(ql); q = q[:p] + [rnd.randrange(4) for _ in range(rnd.randrange(1, 6))] + q[p:]; q = q[:ql]
        args = dict(q=q, t=t, o_del=6, e_del=1, o_ins=rnd.choice([6, 4]), e_ins=rnd.choice([1, 2]), w=rnd.choice([100, 200, 5]),
                    end_bonus=rnd.choice([5, 30, 0]), zdrop=rnd.choice([100, 20, 0]), h0=rnd.randrange(1, 120))
        out['extend2'].append(dict(args=args, res=ref.ksw_extend2(mat=mat, **args)))
        gargs = dict(q=q, t=t, o_del=6, e_del=1, o_ins=6, e_ins=1, w=max(abs(ql - tl) + 3, rnd.choice([3, 10, 50])))
        out['global2'].append(dict(args=gargs, res=ref.ksw_global2(mat=mat, **gargs)))
    with gzip.GzipFile(os.path.join(HERE, 'primitives.json.gz'), 'wb', mtime=0) as g:
        g.write(json.dumps(out).encode())
    print({k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
