"""The seeded workload generator (bsbolt_b200/simulate.py) against the conventions of the reference generator
(`bsbolt Simulate`: bsbolt/Simulate/SimulateMethylatedReads.py:127-178 + the forked wgsim): the same checker reads the
truth line of (a) reads the REFERENCE generator made from its own test genome (tests/golden/c1/se100_1.fq.gz, committed with
the script that made it, tests/golden/make_c1c3_golden.py) and (b) reads of this repo's generator on the same genome, and
verifies what an aligner test relies on: name and truth-line format, 0-based half-open window of each mate, orientation
(which label is printed reverse-complemented) and the direction of the bisulfite conversion per label."""
import gzip
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN

COMP = str.maketrans('ACGTN', 'TGCAN')
NAME = re.compile(r'^@([0-9a-f]+)_([^/\s]+)/([12])$')   # the reference's wgsim prints the id in hexadecimal, this repo's generator in decimal
TRUTH = re.compile(r'^\+([^:]+):(\d+):(\d+):([^:]+):([WC])(C2T|G2A)$')


def genome():
    seqs, name = {}, None
    for line in gzip.open(os.path.join(GOLDEN, 'c1', 'BSB_test.fa.gz'), 'rt'):
        if line.startswith('>'):
            name = line[1:].split()[0]; seqs[name] = []
        else:
            seqs[name].append(line.strip().upper())
    return {k: ''.join(v) for k, v in seqs.items()}


def records(path, limit):
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'rt') as f:
        for _ in range(limit):
            rec = [f.readline().rstrip('\n') for _ in range(4)]
            if not rec[0]:
                return
            yield rec


def check_record(rec, ref, mate):
    """Returns (positions checked, unlabelled positions that differ, labelled positions that break the convention). Per-base truth
    letters of the reference generator: y/c unmethylated CH/CG (converted), Y/C methylated (kept), M anything else that was not
    flagged -- wgsim's sequencing errors (-e 0.005) at bases that cannot be methylated stay 'M' -- other letters = variant,
    error at a methylatable base or indel (skipped)."""
    name, seq, plus, qual = rec
    m, t = NAME.match(name), TRUTH.match(plus)
    assert m and t, (name, plus)
    assert int(m.group(3)) == mate and m.group(2) == t.group(1)          # the contig of the name is the contig of the truth line
    contig, start, end, cigar, strand, conv = t.group(1), int(t.group(2)), int(t.group(3)), t.group(4), t.group(5), t.group(6)
    assert len(seq) == len(qual) and 0 <= start < end <= len(ref[contig])
    window = ref[contig][start:end]
    # a Watson C2T mate and a Crick G2A mate are printed as the forward strand reads; the other two reverse-complemented
    x = window if (strand == 'W') == (conv == 'C2T') else window.translate(COMP)[::-1]
    frm, to = ('C', 'T') if conv == 'C2T' else ('G', 'A')
    per_base = cigar if not cigar[:-1].isdigit() else None
    if per_base is not None and (len(per_base) != len(seq) or len(seq) != end - start or re.search(r'[^MyYcC]', per_base) and re.search(r'\d', per_base)):
        return 0, 0, 0                                                    # an indel shifts the columns: not checked
    if len(seq) != end - start:
        return 0, 0, 0
    n = bad = bad_label = 0
    for i, (s, g) in enumerate(zip(seq, x)):
        c = per_base[i] if per_base is not None else None
        if c is not None and c not in 'MyYcC':
            continue
        n += 1
        if c in ('y', 'c'):
            bad_label += not (g == frm and s == to)
        elif c in ('Y', 'C'):
            bad_label += not (g == frm and s == frm)
        elif c == 'M':
            bad += s != g
        else:                                                             # no per-base truth: the base itself or its conversion
            bad += not (s == g or (g == frm and s == to))
    return n, bad, bad_label


def test_reference_generator_reads_follow_the_checked_conventions():
    ref = genome()
    n = bad = bad_label = recs = 0
    labels = set()
    for rec in records(os.path.join(GOLDEN, 'c1', 'se100_1.fq.gz'), 4000):
        a, b, c = check_record(rec, ref, 1)
        n += a; bad += b; bad_label += c; recs += 1
        labels.add(TRUTH.match(rec[2]).group(5) + TRUTH.match(rec[2]).group(6))
    assert recs == 4000 and n > 300000 and bad_label == 0
    assert bad < 0.005 * n                                               # unflagged sequencing errors only
    assert labels == {'WC2T', 'CC2T'}                                     # directional single-end: always the C2T mate


@pytest.mark.parametrize('paired,undirectional', [(False, False), (True, False), (True, True), (False, True)])
def test_seeded_generator_follows_the_reference_conventions(tmp_path, paired, undirectional):
    from bsbolt_b200 import simulate
    ref = genome()
    names = list(ref)
    contigs = [np.frombuffer(ref[k].encode(), dtype=np.uint8) for k in names]
    paths, n_out = simulate.simulate_reads(names, contigs, str(tmp_path / 'sim'), 3000, read_len=100, paired=paired, undirectional=undirectional,
                                           seed=5, mut_rate=0.0, seq_err=0.0)
    assert n_out == 3000
    labels = [set(), set()]
    for k, p in enumerate(paths):
        n = bad = 0
        for rec in records(p, 3000):
            a, b, c = check_record(rec, ref, k + 1)
            n += a; bad += b + c
            labels[k].add(TRUTH.match(rec[2]).group(5) + TRUTH.match(rec[2]).group(6))
            assert rec[3] == '?' * 99 + '>'                                # the reference's wgsim prints this quality string too
        assert n == 300000 and bad == 0
    if not undirectional:
        assert labels[0] == {'WC2T', 'CC2T'} and (not paired or labels[1] == {'WG2A', 'CG2A'})
    else:
        assert labels[0] == {'WC2T', 'CC2T', 'WG2A', 'CG2A'}
    if paired:   # mates of a pair: same id, same contig, opposite conversion, windows of one fragment
        for r1, r2 in zip(records(paths[0], 500), records(paths[1], 500)):
            assert NAME.match(r1[0]).group(1, 2) == NAME.match(r2[0]).group(1, 2)
            t1, t2 = TRUTH.match(r1[2]), TRUTH.match(r2[2])
            assert t1.group(5) == t2.group(5) and t1.group(6) != t2.group(6)
            assert abs(int(t1.group(2)) - int(t2.group(2))) <= 400
