"""BAM output (`bsbolt Align -O`, the `| stream_bam` half of the reference pipeline): the uncompressed BAM stream of
the native writer (bsbolt_b200/csrc/host_bam.cpp, through the C ABI) must be byte-identical to what the reference's
htslib encoder makes of the same SAM text -- against committed digests of reference runs (tests/golden/bam_golden.json,
made by tests/golden/make_bam_golden.py) and, where oracle/_ref/stream_bam exists, against a live run of it."""
import gzip
import hashlib
import json
import os
import struct
import subprocess

import pytest

from conftest import GOLDEN, ROOT

STREAM_BAM = os.path.join(ROOT, 'oracle', '_ref', 'stream_bam')
EOF_BLOCK = bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')


def golden_streams():
    return json.load(open(os.path.join(GOLDEN, 'bam_golden.json')))['streams']


def sam_bytes(name):
    p = os.path.join(GOLDEN, name)
    return gzip.open(p, 'rb').read() if name.endswith('.gz') else open(p, 'rb').read()


def bgzf_blocks(data):
    """walks the BGZF framing; returns the uncompressed payloads"""
    out, off = [], 0
    while off < len(data):
        assert data[off:off + 4] == b'\x1f\x8b\x08\x04' and data[off + 10:off + 16] == b'\x06\x00BC\x02\x00', f'bad BGZF header at {off}'
        bsize, = struct.unpack('<H', data[off + 16:off + 18])
        block = data[off:off + bsize + 1]
        crc, isize = struct.unpack('<II', block[-8:])
        import zlib
        payload = zlib.decompress(block[18:-8], -15)
        assert len(payload) == isize and isize <= 0xff00 and (zlib.crc32(payload) & 0xffffffff) == crc
        out.append(payload)
        off += bsize + 1
    return out


def walk_records(raw):
    """(header text, reference names, list of record byte strings) of an uncompressed BAM stream"""
    assert raw[:4] == b'BAM\x01'
    l_text, = struct.unpack('<i', raw[4:8])
    text = raw[8:8 + l_text]
    off = 8 + l_text
    n_ref, = struct.unpack('<i', raw[off:off + 4]); off += 4
    names = []
    for _ in range(n_ref):
        l_name, = struct.unpack('<i', raw[off:off + 4]); off += 4
        names.append(raw[off:off + l_name - 1].decode()); off += l_name + 4
    recs = []
    while off < len(raw):
        bs, = struct.unpack('<i', raw[off:off + 4])
        recs.append(raw[off:off + 4 + bs]); off += 4 + bs
    assert off == len(raw)
    return text, names, recs


@pytest.mark.parametrize('name', sorted(json.load(open(os.path.join(GOLDEN, 'bam_golden.json')))['streams']))
@pytest.mark.parametrize('threads', [1, 3])
def test_stream_identical_to_reference_encoder(built, tmp_path, name, threads):
    from bsbolt_b200 import _native
    want = golden_streams()[name]
    src = tmp_path / 'in.sam'
    src.write_bytes(sam_bytes(name))
    fd = os.open(src, os.O_RDONLY)
    try:
        n = _native.stream_bam(fd, tmp_path / 'out.bam', threads=threads, level=-1)
    finally:
        os.close(fd)
    data = open(tmp_path / 'out.bam', 'rb').read()
    assert data.endswith(EOF_BLOCK)
    blocks = bgzf_blocks(data)
    raw = b''.join(blocks)
    assert len(raw) == want['raw_len'] and hashlib.sha256(raw).hexdigest() == want['raw_sha256']
    assert gzip.open(tmp_path / 'out.bam', 'rb').read() == raw            # also readable as a plain multi-member gzip file
    text, names, recs = walk_records(raw)
    assert n == len(recs) == sum(1 for l in sam_bytes(name).split(b'\n') if l and not l.startswith(b'@'))
    # like htslib, no record is split across blocks unless it is larger than a block
    ends, pos = set(), 0
    for b in blocks:
        pos += len(b); ends.add(pos)
    off = len(raw) - sum(len(r) for r in recs)
    for r in recs:
        if len(r) <= 0xff00:
            inside = [e for e in ends if off < e < off + len(r)]
            assert not inside, 'a record that fits a block was split'
        off += len(r)
    if os.path.exists(STREAM_BAM):   # the reference encoder itself, live
        subprocess.run([STREAM_BAM, '-o', str(tmp_path / 'ref.bam')], stdin=open(src), check=True, stderr=subprocess.DEVNULL)
        assert gzip.open(tmp_path / 'ref.bam', 'rb').read() == raw


def test_levels_and_empty_input(built, tmp_path):
    from bsbolt_b200 import _native
    src = tmp_path / 'in.sam'
    src.write_bytes(sam_bytes('se100.sam.gz'))
    raws = []
    for level in (0, 1, 9):
        fd = os.open(src, os.O_RDONLY)
        _native.stream_bam(fd, tmp_path / f'l{level}.bam', threads=2, level=level)
        os.close(fd)
        raws.append(b''.join(bgzf_blocks(open(tmp_path / f'l{level}.bam', 'rb').read())))
    assert raws[0] == raws[1] == raws[2]
    assert os.path.getsize(tmp_path / 'l9.bam') < os.path.getsize(tmp_path / 'l1.bam') < os.path.getsize(tmp_path / 'l0.bam')
    # header only, and nothing at all
    for body in (b'@SQ\tSN:c\tLN:5\n', b''):
        src.write_bytes(body)
        fd = os.open(src, os.O_RDONLY)
        assert _native.stream_bam(fd, tmp_path / 'e.bam', threads=1) == 0
        os.close(fd)
        text, names, recs = walk_records(gzip.open(tmp_path / 'e.bam', 'rb').read())
        assert text == body and recs == [] and names == (['c'] if body else [])


def test_malformed_records_fail_loudly(built, tmp_path):
    from bsbolt_b200 import _native
    src = tmp_path / 'in.sam'
    for bad in (b'r\t0\tc\t1\t0\t4M\t*\t0\t0\tACG\tIII\n',        # CIGAR/SEQ length mismatch
                b'r\t0\tc\t1\t0\t3M\t*\t0\t0\tACG\tII\n',         # SEQ/QUAL length mismatch
                b'r\t0\tc\t1\t0\t3Q\t*\t0\t0\tACG\tIII\n',        # unknown CIGAR operator
                b'r\t0\tc\t1\n'):                                  # truncated
        src.write_bytes(b'@SQ\tSN:c\tLN:50\n' + bad)
        fd = os.open(src, os.O_RDONLY)
        with pytest.raises(RuntimeError, match='bam_encode'):
            _native.stream_bam(fd, tmp_path / 'x.bam', threads=1)
        os.close(fd)


@pytest.mark.gpu
def test_aligner_writes_the_bam_the_reference_pipeline_would(built, golden, tmp_path):
    """bsb_mem_main_bam == bsb_mem_main | stream_bam: same records as the golden SAM, identical uncompressed stream."""
    from bsbolt_b200 import _native
    from conftest import strip_pg
    idx = _native.Index(golden.idxbase, 0)
    try:
        for case in ('pe150', 'se100_un'):
            argv = golden.argv(case) + []
            argv = argv[:-len(golden.cases[case]['fq']) - 1] + ['-K', '150000'] + argv[-len(golden.cases[case]['fq']) - 1:]
            with open(tmp_path / 'log', 'w') as fl:
                rc, st = _native.mem_main_bam(argv, tmp_path / f'{case}.bam', index=idx, threads=3, level=1, log_fd=fl.fileno())
            assert rc == 0, _native.last_error()
            with open(tmp_path / f'{case}.sam', 'w') as fo, open(tmp_path / 'log2', 'w') as fl:
                rc, _ = _native.mem_main(argv, index=idx, out_fd=fo.fileno(), log_fd=fl.fileno())
            assert rc == 0
            bsstat = lambda f: [l for l in open(tmp_path / f) if l.startswith('BSStat ')]
            assert bsstat('log') == bsstat('log2') and len(bsstat('log')) >= 8          # same BSStat lines
            sam = open(tmp_path / f'{case}.sam').read()
            assert strip_pg(sam) == golden.sam(case)
            raw = b''.join(bgzf_blocks(open(tmp_path / f'{case}.bam', 'rb').read()))
            fd = os.open(tmp_path / f'{case}.sam', os.O_RDONLY)
            _native.stream_bam(fd, tmp_path / 'via_sam.bam', threads=1)
            os.close(fd)
            assert gzip.open(tmp_path / 'via_sam.bam', 'rb').read() == raw
            if os.path.exists(STREAM_BAM):
                subprocess.run([STREAM_BAM, '-o', str(tmp_path / 'ref.bam')], stdin=open(tmp_path / f'{case}.sam'), check=True, stderr=subprocess.DEVNULL)
                assert gzip.open(tmp_path / 'ref.bam', 'rb').read() == raw
            assert len(walk_records(raw)[2]) == golden.cases[case]['n_records']
    finally:
        idx.close()


# ----------------------------------------------------------------------------------------------------------------------
# BAM made on the device (bsb_bam.h records + arbiter, bsb_deflate.h BGZF blocks)
# ----------------------------------------------------------------------------------------------------------------------
BAMSIM = os.path.join(ROOT, 'tests', 'hostsim', 'bamsim')
HOSTSIM = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim')


@pytest.mark.parametrize('order', ['', 'reverse', 'shuffle'])
def test_device_deflate_blocks_inflate_to_the_input(built, tmp_path, order):
    """bgzf_block (the body of k_bgzf_deflate) phase by phase on the CPU: every block is valid deflate for zlib, carries
    zlib's CRC-32, and the bytes do not depend on the order the block's threads run in"""
    import random
    random.seed(7)
    cases = {
        'empty': b'', 'one': b'x', 'three': b'abc', 'four': b'abcd', 'zeros': bytes(200000), 'noise': random.randbytes(150000),
        'runs': b''.join(bytes([random.randrange(4)]) * random.randrange(1, 600) for _ in range(1500)),
        'records': b''.join(b'read_%d\tNM:i:%d\tMD:Z:%d\tYS:Z:W_C2T\t%s\n' % (i, i % 7, i % 150, bytes(random.choices(b'FFFFFF:,#', k=40))) for i in range(20000)),
        'block': bytes(range(256)) * 255, 'block+1': bytes(range(256)) * 255 + b'z',
        'skewed': bytes(random.choices(range(256), weights=[2 ** (-(i / 4)) for i in range(256)], k=200000)),
    }
    # seeded mixtures: literals, runs, near and far repeats (up to the 32 KB window and beyond), lengths around the block and chunk edges
    for k in range(24):
        parts, total = [], random.choice([5, 511, 512, 513, 4097, 65279, 65280, 65281, 130560, random.randrange(1, 200000)])
        while sum(map(len, parts)) < total:
            kind = random.random()
            if kind < 0.3 or not parts:
                parts.append(random.randbytes(random.randrange(1, 300)))
            elif kind < 0.5:
                parts.append(bytes([random.randrange(256)]) * random.randrange(1, 700))
            else:
                flat = b''.join(parts)
                a = random.randrange(len(flat)); l = random.randrange(3, 400)
                if random.random() < 0.3:
                    a = max(0, len(flat) - random.randrange(1, 70000))
                parts.append(flat[a:a + l])
        cases[f'mix{k}'] = b''.join(parts)[:total]
    env = dict(os.environ)
    if order:
        env['BSB_PAR_ORDER'] = order
    sizes = {}
    for name, data in cases.items():
        (tmp_path / 'in').write_bytes(data)
        p = subprocess.run([BAMSIM, 'deflate', str(tmp_path / 'in'), str(tmp_path / 'out')], env=env, capture_output=True, text=True)
        assert p.returncode == 0, (name, p.stderr)
        out = (tmp_path / 'out').read_bytes()
        assert out.endswith(EOF_BLOCK)
        assert b''.join(bgzf_blocks(out)) == data, name
        sizes[name] = hashlib.sha256(out).hexdigest()
        if name in ('zeros', 'runs', 'records', 'block'):
            assert len(out) < len(data) / (3 if name == 'records' else 20), name   # the matches are found
        if name == 'noise':
            assert len(out) <= len(data) + 40 * (len(data) // 0xff00 + 2)   # stored blocks
    ref = tmp_path.parent / 'deflate_digests.json'                 # shared by the three parametrisations of this session
    if ref.exists():
        assert json.load(open(ref)) == sizes, 'the compressed bytes depend on the thread order'
    else:
        json.dump(sizes, open(ref, 'w'))


@pytest.mark.parametrize('name', sorted(json.load(open(os.path.join(GOLDEN, 'bam_golden.json')))['streams']))
def test_device_record_encoder_identical_to_reference_encoder(built, tmp_path, name):
    """bam_entry with the sorted-contig lookup (the body of k_bam_count / k_bam_write) + bgzf_block on the reference's SAM:
    the uncompressed stream equals the reference encoder's, and no record that fits a block is split"""
    want = golden_streams()[name]
    src = tmp_path / 'in.sam'
    src.write_bytes(sam_bytes(name))
    p = subprocess.run([BAMSIM, 'sam2bam', str(src), str(tmp_path / 'out.bam')], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    data = open(tmp_path / 'out.bam', 'rb').read()
    blocks = bgzf_blocks(data)
    raw = b''.join(blocks)
    assert len(raw) == want['raw_len'] and hashlib.sha256(raw).hexdigest() == want['raw_sha256']
    text, names, recs = walk_records(raw)
    ends, pos = set(), 0
    for b in blocks:
        pos += len(b); ends.add(pos)
    off = len(raw) - sum(len(r) for r in recs)
    for r in recs:
        if len(r) <= 0xff00 // 2:
            assert not [e for e in ends if off < e < off + len(r)], 'a record that fits a block was split'
        off += len(r)


@pytest.mark.parametrize('case', ['pe150', 'pe150_un', 'pe150_un_sp0', 'se100_un', 'se50_clip'])
def test_device_bam_stage_equals_host_writer(built, golden, tmp_path, case):
    """The whole stage on the CPU (tests/hostsim, HOSTSIM_BAM): arbiter on "device", records, blocks -- against the host
    path (SAM text -> arbiter -> BamWriter) of the same run: identical uncompressed stream, identical BSStat lines. The
    undirectional cases drop the losing conversion group and rewrite BS-ambiguous reads as unmapped records."""
    outs, stats = [], []
    for host in (False, True):
        env = dict(os.environ, BSB_HOSTSIM_SEED_V3='1', HOSTSIM_BAM=str(tmp_path / f'o{int(host)}.bam'))
        if host:
            env['HOSTSIM_BAM_HOST'] = '1'
        argv = golden.argv(case)
        nfq = len(golden.cases[case]['fq'])
        argv = argv[:-nfq - 1] + ['-K', '60000'] + argv[-nfq - 1:]
        p = subprocess.run([HOSTSIM] + argv, capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(b''.join(bgzf_blocks(open(tmp_path / f'o{int(host)}.bam', 'rb').read())))
        stats.append([l for l in p.stderr.split('\n') if l.startswith('BSStat ')])
    assert outs[0] == outs[1]
    assert stats[0] == stats[1] and len(stats[0]) >= 8
    assert len(walk_records(outs[0])[2]) == golden.cases[case]['n_records']


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['pe150', 'pe150_un', 'se100_un', 'se50_clip'])
def test_bam_made_on_the_device_equals_the_host_writer(built, golden, tmp_path, case):
    """bsb_mem_main_bam at the default level: arbiter (k_bam_arbiter), records (k_bam_count / k_bam_write) and BGZF blocks
    (k_bgzf_deflate) on the GPU -- against the same run through SAM text and the host encoder: identical uncompressed
    stream (which the other tests pin to the reference's stream_bam), identical BSStat lines, every block valid for zlib."""
    from bsbolt_b200 import _native
    idx = _native.Index(golden.idxbase, 0)
    try:
        argv = golden.argv(case)
        nfq = len(golden.cases[case]['fq'])
        argv = argv[:-nfq - 1] + ['-K', '60000'] + argv[-nfq - 1:]
        raws, logs, stats = [], [], []
        for level in (-1, 1):
            with open(tmp_path / f'log{level}', 'w') as fl:
                rc, st = _native.mem_main_bam(argv, tmp_path / f'l{level}.bam', index=idx, threads=2, level=level, log_fd=fl.fileno())
            assert rc == 0, _native.last_error()
            data = open(tmp_path / f'l{level}.bam', 'rb').read()
            assert data.endswith(EOF_BLOCK)
            raws.append(b''.join(bgzf_blocks(data)))
            logs.append([l for l in open(tmp_path / f'log{level}') if l.startswith('BSStat ')])
            stats.append(st)
        assert raws[0] == raws[1]
        assert logs[0] == logs[1] and len(logs[0]) >= 8
        assert stats[0]['d2h_bytes'] < stats[1]['d2h_bytes']           # compressed blocks crossed PCIe, not text
        assert len(walk_records(raws[0])[2]) == golden.cases[case]['n_records']
        if os.path.exists(STREAM_BAM):
            with open(tmp_path / 'x.sam', 'w') as fo, open(tmp_path / 'log2', 'w') as fl:
                rc, _ = _native.mem_main(argv, index=idx, out_fd=fo.fileno(), log_fd=fl.fileno())
            subprocess.run([STREAM_BAM, '-o', str(tmp_path / 'ref.bam')], stdin=open(tmp_path / 'x.sam'), check=True, stderr=subprocess.DEVNULL)
            assert gzip.open(tmp_path / 'ref.bam', 'rb').read() == raws[0]
    finally:
        idx.close()
