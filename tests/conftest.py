import gzip
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _gunzip(src, dst):
    with gzip.open(src, 'rb') as f, open(dst, 'wb') as o:
        shutil.copyfileobj(f, o)


@pytest.fixture(scope='session')
def golden(tmp_path_factory):
    """Unpacks tests/golden into a scratch directory: index files, FASTQs, manifest."""
    d = tmp_path_factory.mktemp('golden')
    os.makedirs(d / 'db')
    for f in os.listdir(os.path.join(GOLDEN, 'db')):
        _gunzip(os.path.join(GOLDEN, 'db', f), d / 'db' / f[:-3])
    for f in os.listdir(GOLDEN):
        if f.endswith('.fq.gz') or f == 'genome.fa.gz':
            _gunzip(os.path.join(GOLDEN, f), d / f[:-3])
    # the database directory of `bsbolt Index` also holds the FASTA it indexed (one line per contig)
    with open(d / 'genome.fa') as f, open(d / 'db' / 'BSB_ref.fa', 'w') as o:
        first = True
        for line in f:
            if line.startswith('>'):
                o.write(('' if first else '\n') + line.split()[0] + '\n'); first = False
            else:
                o.write(line.strip())
        o.write('\n')
    man = json.load(open(os.path.join(GOLDEN, 'golden.json')))

    class G:
        dir = str(d)
        idxbase = str(d / 'db' / 'BSB_ref.fa')
        manifest = man
        cases = man['cases']

        @staticmethod
        def argv(case):
            c = man['cases'][case]
            return ['mem'] + man['launcher_args'] + c['extra'] + [G.idxbase] + [str(d / f) for f in c['fq']]

        @staticmethod
        def sam(case):
            return gzip.open(os.path.join(GOLDEN, case + '.sam.gz'), 'rt').read()
    return G


def strip_pg(text):
    return ''.join(l + '\n' for l in text.split('\n') if l and not l.startswith('@PG'))


def first_diff(a, b):
    la, lb = a.split('\n'), b.split('\n')
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f'line {i}:\n  expected: {x[:300]}\n  got:      {y[:300]}'
    return f'length differs: {len(la)} vs {len(lb)} lines'


@pytest.fixture(scope='session')
def built():
    """Builds the product library, the oracle and the CPU unit harness once per session (no GPU needed)."""
    import __graft_entry__ as ge
    ge.build()
    return ge


def run_capture(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)
