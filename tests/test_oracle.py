"""Pins the oracle restatement (oracle/bsb_oracle.c): (1) against known-answer vectors recorded from
the reference's own functions (tests/golden/primitives.json.gz), always; (2) against the live reference
library on fresh random inputs when oracle/_ref/libbwa_ref.so is present."""
import gzip
import json
import os
import random

import pytest

from conftest import ROOT
from reflib import GOLDEN, OracleLib, RefLib, REF_SO, unpack_index


@pytest.fixture(scope='module')
def libs(built, tmp_path_factory):
    d = unpack_index(str(tmp_path_factory.mktemp('idx')))
    base = os.path.join(d, 'BSB_ref.fa')
    return OracleLib(base), (RefLib(base) if os.path.exists(REF_SO) else None)


@pytest.fixture(scope='module')
def vectors():
    return json.loads(gzip.open(os.path.join(GOLDEN, 'primitives.json.gz'), 'rt').read())


def test_occ4_extend_known_answers(libs, vectors):
    o, _ = libs
    for k, want in vectors['occ4']:
        assert o.occ4(k) == want
    for v in vectors['extend']:
        assert o.extend(v['ik'], v['is_back']) == v['ok']


def test_smem_known_answers(libs, vectors):
    o, _ = libs
    for v in vectors['smem']:
        assert o.smem1(v['q'], v['x'], v['min_intv']) == v['res']
    for v in vectors['seed_forward']:
        assert o.seed_strategy1(v['q'], v['x'], 19, 20) == v['res']


def test_sa_known_answers(libs, vectors):
    o, _ = libs
    for k, want in vectors['sa']:
        assert o.sa(k) == want


def test_dp_known_answers(libs, vectors):
    o, _ = libs
    mat = o.scmat(1, 4)
    for v in vectors['extend2']:
        assert o.ksw_extend2(mat=mat, **v['args']) == v['res']
    for v in vectors['global2']:
        assert o.ksw_global2(mat=mat, **v['args']) == v['res']


def test_oracle_vs_live_reference(libs):
    o, r = libs
    if r is None:
        pytest.skip('oracle/_ref/libbwa_ref.so not present')
    rnd = random.Random(99)
    for _ in range(400):
        k = rnd.randrange(0, o.seq_len + 1)
        assert o.occ4(k) == r.occ4(k)
        if k:
            assert o.sa(k) == r.sa(k)
    mat = o.scmat(1, 4)
    for _ in range(200):
        ql, tl = rnd.randrange(1, 151), rnd.randrange(1, 200)
        t = [rnd.randrange(4) for _ in range(tl)]
        q = [(t[i] if i < tl and rnd.random() > 0.1 else rnd.randrange(5)) for i in range(ql)]
        a = dict(q=q, t=t, o_del=6, e_del=1, o_ins=6, e_ins=1, w=rnd.choice([100, 30]), end_bonus=rnd.choice([30, 5]),
                 zdrop=100, h0=rnd.randrange(1, 150))
        assert o.ksw_extend2(mat=mat, **a) == r.ksw_extend2(mat=mat, **a)
        g = dict(q=q, t=t, o_del=6, e_del=1, o_ins=6, e_ins=1, w=abs(ql - tl) + rnd.choice([3, 20]))
        assert o.ksw_global2(mat=mat, **g) == r.ksw_global2(mat=mat, **g)
    genome = ''.join(l.strip() for l in gzip.open(os.path.join(GOLDEN, 'genome.fa.gz'), 'rt') if not l.startswith('>')).upper()
    for _ in range(60):
        p = rnd.randrange(0, len(genome) - 150)
        s = genome[p:p + rnd.randrange(30, 150)].replace('C', 'T')
        q = [o.code(c) for c in s]
        for _ in range(rnd.randrange(0, 4)):
            q[rnd.randrange(len(q))] = rnd.randrange(5)
        x = rnd.randrange(len(q))
        assert o.smem1(q, x, 1) == r.smem1(q, x, 1)
        assert o.seed_strategy1(q, x, 19, 20) == r.seed_strategy1(q, x, 19, 20)
