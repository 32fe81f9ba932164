"""Interleaved input for the smart-pairing (`-p`) cases, derived deterministically from the committed FASTQ
fixtures: pairs of pe150c_1/2 back to back, a single-end read of se100c every 7th pair, an orphaned first mate
every 11th. Shared by tests/golden/make_smart_golden.py (which ran the reference on it) and the tests."""
import os


def _records(path):
    with open(path) as f:
        while True:
            r = [f.readline() for _ in range(4)]
            if not r[0]:
                return
            yield ''.join(r)


def write_interleaved(fq_dir, out_path, n_pairs=1500):
    se = _records(os.path.join(fq_dir, 'se100c.fq'))
    with open(out_path, 'w') as out:
        pairs = zip(_records(os.path.join(fq_dir, 'pe150c_1.fq')), _records(os.path.join(fq_dir, 'pe150c_2.fq')))
        for n, (a, b) in enumerate(pairs, 1):
            if n > n_pairs:
                break
            if n % 7 == 3:
                out.write(next(se))
            out.write(a)
            if n % 11 != 5:
                out.write(b)
    return str(out_path)


SMART_CASES = {'smart_p': ['-p', '-K', '100000'], 'smart_p_un': ['-z', '-p', '-K', '70000']}
