"""Seeded, vectorised workload generator for benchmarks and differential tests.

Follows the conventions of the reference generator (`bsbolt Simulate`: bsbolt/Simulate/SimulateMethylatedReads.py
+ External/WGSIM/src/wgsim.cpp) -- read names `@<id>_<contig>/1|2`, truth on the `+` line as
`contig:start:end:cigar:{W|C}{C2T|G2A}` (start/end: the mate's own forward-strand window, 0-based, end exclusive; the cigar is
`<L>M` here where the reference prints one letter per base), checked against reads of the reference generator in
tests/test_simulator_conventions.py, directional reads C->T on read 1 / G->A on read 2, optional
undirectional swap, SNP/indel/sequencing-error rates with the same defaults -- but is reproducible from its
seed (the reference mixes std::random_device into wgsim and unseeded Python `random`) and fast enough to
produce millions of pairs inside a benchmark run. It is NOT on the alignment path.
"""
import gzip
import os

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b'ACGTN', b'TGCAN'):
    _COMP[_a] = _b


def _cname(i, n):
    return 'chr%0*d' % (len(str(n)), i + 1)


def make_genome(path, contig_lens, seed=20240517, n_frac=0.005, n_dups=4, dup_len=5000):
    """i.i.d. ACGT contigs with a few N runs and duplicated segments (so XA/YC/MAPQ-0 paths fire)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b'ACGT', dtype=np.uint8)
    contigs = []
    for n in contig_lens:
        s = alphabet[rng.integers(0, 4, size=n, dtype=np.uint8)]
        k = int(n * n_frac)
        while k > 0:
            run = int(min(k, rng.integers(20, 400)))
            p = int(rng.integers(0, max(1, n - run)))
            s[p:p + run] = ord('N')
            k -= run
        contigs.append(s)
    for _ in range(n_dups):
        a, b = rng.integers(0, len(contigs), size=2)
        if len(contigs[a]) <= dup_len or len(contigs[b]) <= dup_len:
            continue
        pa = int(rng.integers(0, len(contigs[a]) - dup_len))
        pb = int(rng.integers(0, len(contigs[b]) - dup_len))
        contigs[b][pb:pb + dup_len] = contigs[a][pa:pa + dup_len]
    with open(path, 'wb') as f:
        for i, s in enumerate(contigs):
            f.write(b'>' + _cname(i, len(contigs)).encode() + b'\n')
            n = len(s)
            full = (n // 60) * 60
            if full:
                block = np.empty((full // 60, 61), dtype=np.uint8)
                block[:, :60] = s[:full].reshape(-1, 60)
                block[:, 60] = 10
                f.write(block.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b'\n')
    return [_cname(i, len(contigs)) for i in range(len(contigs))], contigs


def read_fasta(path):
    names, seqs, cur = [], [], []
    op = gzip.open if str(path).endswith('.gz') else open
    with op(path, 'rb') as f:
        for line in f:
            if line.startswith(b'>'):
                if cur:
                    seqs.append(np.frombuffer(b''.join(cur), dtype=np.uint8))
                names.append(line[1:].split()[0].decode())
                cur = []
            else:
                cur.append(line.strip().upper())
    seqs.append(np.frombuffer(b''.join(cur), dtype=np.uint8))
    return names, seqs


def simulate_reads(names, contigs, out_prefix, n_pairs, read_len=150, paired=True, undirectional=False, seed=7,
                   insert_mean=50, insert_sd=50, mut_rate=0.005, indel_frac=0.2, seq_err=0.001,
                   cpg_meth=0.8, ch_meth=0.02, corrupt_frac=0.0, chunk=200000, first_id=0, truth=True):
    """Writes <out_prefix>_1.fq (and _2.fq). Returns the file paths."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    genome = np.concatenate(contigs)
    offs = np.concatenate([[0], np.cumsum(lens)])
    paths = [f'{out_prefix}_1.fq'] + ([f'{out_prefix}_2.fq'] if paired else [])
    files = [open(p, 'wb') for p in paths]
    L = read_len
    qual = (b'?' * (L - 1) + b'>')
    done = 0
    rid = first_id
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    fast_names = None
    if not truth and len({len(x) for x in names}) == 1:
        fast_names = np.array([np.frombuffer(x.encode(), dtype=np.uint8) for x in names])
    while done < n_pairs:
        n = int(min(chunk, (n_pairs - done) * 1.03 + 64))  # a few candidates are rejected (N-rich, contig end)
        frag = (2 * L + np.clip(rng.normal(insert_mean, insert_sd, size=n), -L + 10, 400).astype(np.int64)) if paired else np.full(n, L, dtype=np.int64)
        ci = rng.choice(len(contigs), size=n, p=lens / lens.sum())
        ok_len = lens[ci] > frag + 2
        start = (rng.random(n) * np.maximum(lens[ci] - frag - 1, 1)).astype(np.int64)
        gpos = offs[ci] + start
        watson = rng.random(n) < 0.5
        idx = np.arange(L, dtype=np.int64)
        # 5' read of the fragment and the far end, both on the genome's forward strand
        left = genome[gpos[:, None] + idx[None, :]]
        right = genome[(gpos + frag - L)[:, None] + idx[None, :]] if paired else left
        nxt_l = genome[np.minimum(gpos[:, None] + idx[None, :] + 1, len(genome) - 1)]
        nxt_r = genome[np.minimum((gpos + frag - L)[:, None] + idx[None, :] + 1, len(genome) - 1)]
        prv_l = genome[np.maximum(gpos[:, None] + idx[None, :] - 1, 0)]
        prv_r = genome[np.maximum((gpos + frag - L)[:, None] + idx[None, :] - 1, 0)]

        def mutate(m):
            sub = rng.random(m.shape) < (mut_rate * (1 - indel_frac) + seq_err)
            m = m.copy()
            m[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
            return m

        def bisulfite(m, nxt, prv, top):
            """top strand: unmethylated C -> T; bottom strand (seen on the forward strand): G -> A"""
            m = m.copy()
            if top is None:
                return m
            r = rng.random(m.shape)
            for is_top in (True, False):
                rows = top if is_top else ~top
                base, conv = (ord('C'), ord('T')) if is_top else (ord('G'), ord('A'))
                cpg = (nxt == ord('G')) if is_top else (prv == ord('C'))
                site = (m == base) & rows[:, None]
                keep = np.where(cpg, r < cpg_meth, r < ch_meth)
                m[site & ~keep] = conv
            return m

        left_m, right_m = mutate(left), mutate(right)
        left_b = bisulfite(left_m, nxt_l, prv_l, watson)
        right_b = bisulfite(right_m, nxt_r, prv_r, watson)
        # watson fragment: R1 = left (C2T), R2 = revcomp(right) (G2A); crick fragment: R1 = revcomp(right), R2 = left
        rc_right = _COMP[right_b[:, ::-1]]
        rc_left = _COMP[left_b[:, ::-1]]
        r1 = np.where(watson[:, None], left_b, rc_right)
        r2 = np.where(watson[:, None], rc_right, left_b)
        if not paired:
            r1 = np.where(watson[:, None], left_b, rc_left)
        swap = (rng.random(n) < 0.5) if undirectional else np.zeros(n, dtype=bool)
        if paired:
            r1, r2 = np.where(swap[:, None], r2, r1), np.where(swap[:, None], r1, r2)
        else:
            r1 = np.where(swap[:, None], _COMP[r1[:, ::-1]], r1)
        # small indels: delete or duplicate one base somewhere in the read (keeps the length fixed)
        for m in ((r1, r2) if paired else (r1,)):
            has = rng.random(n) < mut_rate * indel_frac * L
            rows = np.nonzero(has)[0]
            for i in rows:
                p = int(rng.integers(5, L - 5))
                if rng.random() < 0.5:
                    m[i, p:-1] = m[i, p + 1:].copy()
                else:
                    m[i, p + 1:] = m[i, p:-1].copy()
        if corrupt_frac > 0 and paired:  # force mate rescue: 15 % substitutions on a fraction of R2
            rows = rng.random(n) < corrupt_frac
            sub = (rng.random(r2.shape) < 0.15) & rows[:, None]
            r2[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
        has_n = ((r1 == ord('N')).mean(axis=1) > 0.05) | (paired & ((r2 == ord('N')).mean(axis=1) > 0.05))
        good = np.nonzero(ok_len & ~has_n)[0][:n_pairs - done]
        if fast_names is not None:
            # vectorised writer: fixed-width records "@<10-digit id>_<contig>/1\n<seq>\n+\n<qual>\n"
            g = len(good)
            ids = rid + np.arange(g, dtype=np.int64)
            digits = ((ids[:, None] // (10 ** np.arange(9, -1, -1, dtype=np.int64))[None, :]) % 10 + 48).astype(np.uint8)
            cn = fast_names[ci[good]]
            w = cn.shape[1]
            rec_len = 1 + 10 + 1 + w + 3 + L + 3 + L + 1
            for k, (fh, mat) in enumerate(zip(files, (r1, r2) if paired else (r1,))):
                rec = np.empty((g, rec_len), dtype=np.uint8)
                o = 0
                rec[:, o] = ord('@'); o += 1
                rec[:, o:o + 10] = digits; o += 10
                rec[:, o] = ord('_'); o += 1
                rec[:, o:o + w] = cn; o += w
                rec[:, o:o + 3] = np.frombuffer(b'/%d\n' % (k + 1), dtype=np.uint8); o += 3
                rec[:, o:o + L] = mat[good]; o += L
                rec[:, o:o + 3] = np.frombuffer(b'\n+\n', dtype=np.uint8); o += 3
                rec[:, o:o + L] = np.frombuffer(qual, dtype=np.uint8); o += L
                rec[:, o] = 10
                fh.write(rec.tobytes())
            rid += g
        else:
            for j in good:
                c = names[ci[j]]
                tag1 = ('W' if watson[j] else 'C') + ('G2A' if swap[j] else 'C2T')
                s0, e0 = int(start[j]), int(start[j] + frag[j])
                # like the reference generator (SimulateMethylatedReads.output_sim_reads), every mate carries the forward-strand
                # window it was read from: the C2T mate of a Watson fragment and the G2A mate of a Crick fragment the left one
                left_w, right_w = (s0, s0 + L), (e0 - L, e0)
                w1 = left_w if bool(watson[j]) != bool(swap[j]) else right_w
                files[0].write(b'@%d_%s/1\n' % (rid, c.encode()) + r1[j].tobytes() + b'\n+%s:%d:%d:%dM:%s\n' % (c.encode(), w1[0], w1[1], L, tag1.encode()) + qual + b'\n')
                if paired:
                    tag2 = ('W' if watson[j] else 'C') + ('C2T' if swap[j] else 'G2A')
                    w2 = right_w if w1 is left_w else left_w
                    files[1].write(b'@%d_%s/2\n' % (rid, c.encode()) + r2[j].tobytes() + b'\n+%s:%d:%d:%dM:%s\n' % (c.encode(), w2[0], w2[1], L, tag2.encode()) + qual + b'\n')
                rid += 1
        done += len(good)
    for f in files:
        f.close()
    return paths, rid - first_id


def simulate_rrbs_reads(names, contigs, out_path, n_reads, read_len=50, seed=7, lower_bound=30, upper_bound=500,
                        site=b'CCGG', cut_offset=1, mut_rate=0.005, seq_err=0.001, cpg_meth=0.8, ch_meth=0.02, undirectional=False):
    """Single-end reads of a reduced-representation library (BASELINE config C3): every read starts at the 5' end of
    one strand of a restriction fragment (MspI, C^CGG) whose length lies in the size window of `bsbolt Index -rrbs`
    (bsbolt/Index/RRBSIndex.py:103-139), so the reads pile up on a small set of start positions -- the high-duplicate,
    short-read regime -- and fall inside the unmasked part of an RRBS database. Directional library: both strands
    are read 5'->3' with unmethylated C -> T; `undirectional` also emits the reverse complements."""
    rng = np.random.default_rng(seed)
    L = read_len
    starts = []          # (contig, forward-strand start of the read window, is_bottom_strand)
    pat = np.frombuffer(site, dtype=np.uint8)
    for ci, s in enumerate(contigs):
        hit = np.ones(len(s) - len(pat) + 1, dtype=bool)
        for k, c in enumerate(pat):
            hit &= s[k:len(s) - len(pat) + 1 + k] == c
        cuts = np.nonzero(hit)[0] + cut_offset
        frag = np.diff(cuts)
        ok = np.nonzero((frag >= max(lower_bound, L)) & (frag <= upper_bound))[0]
        for k in ok:
            starts.append((ci, int(cuts[k]), 0))                                   # top strand, left end of the fragment
            starts.append((ci, int(cuts[k + 1]) + len(site) - 2 * cut_offset - L, 1))   # bottom strand, right end
    starts = np.array(starts, dtype=np.int64)
    pick = starts[rng.integers(0, len(starts), size=n_reads)]
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    idx = np.arange(L, dtype=np.int64)
    qual = b'?' * (L - 1) + b'>'
    with open(out_path, 'wb') as f:
        for ci in range(len(contigs)):
            rows = pick[pick[:, 0] == ci]
            if not len(rows):
                continue
            s = contigs[ci]
            pos = np.clip(rows[:, 1], 0, len(s) - L)
            m = s[pos[:, None] + idx[None, :]].copy()
            nxt = s[np.minimum(pos[:, None] + idx[None, :] + 1, len(s) - 1)]
            prv = s[np.maximum(pos[:, None] + idx[None, :] - 1, 0)]
            sub = rng.random(m.shape) < (mut_rate + seq_err)
            m[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
            bottom = rows[:, 2] == 1
            r = rng.random(m.shape)
            for is_top in (True, False):
                sel = ~bottom if is_top else bottom
                base, conv = (ord('C'), ord('T')) if is_top else (ord('G'), ord('A'))
                cpg = (nxt == ord('G')) if is_top else (prv == ord('C'))
                keep = np.where(cpg, r < cpg_meth, r < ch_meth)
                m[(m == base) & sel[:, None] & ~keep] = conv
            m = np.where(bottom[:, None], _COMP[m[:, ::-1]], m)
            if undirectional:
                flip = rng.random(len(m)) < 0.5
                m = np.where(flip[:, None], _COMP[m[:, ::-1]], m)
            for j in range(len(m)):
                f.write(b'@rrbs%d_%s_%d_%d\n' % (ci, names[ci].encode(), int(pos[j]), j) + m[j].tobytes() + b'\n+\n' + qual + b'\n')
    return out_path, n_reads
