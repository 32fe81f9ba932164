"""SAM -> BAM for `-O` (the reference pipes into htslib's stream_bam, External/HTSLIB/stream_bam.c).

Host-side glue, not on the hot path: the SAM stream of the aligner is written to a pipe and encoded to
BGZF/BAM by a small pure-Python encoder running in a thread pool of `output_threads` deflaters.
"""
import os
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor

from bsbolt_b200 import _native

_CIG = {c: i for i, c in enumerate('MIDNSHP=X')}
_SEQ = {c: i for i, c in enumerate('=ACMGRSVTWYHKDBN')}
_EOF = bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000')


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_block(data, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b'\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00' + struct.pack('<H', bsize)
            + comp + struct.pack('<II', zlib.crc32(data) & 0xffffffff, len(data)))


def _aux(tag):
    t, ty, v = tag[:2].encode(), tag[3], tag[5:]
    if ty == 'i':
        x = int(v)
        for code, fmt, lo, hi in (('C', '<B', 0, 255), ('c', '<b', -128, 127), ('S', '<H', 0, 65535),
                                  ('s', '<h', -32768, 32767), ('I', '<I', 0, 4294967295), ('i', '<i', -2147483648, 2147483647)):
            if lo <= x <= hi:
                return t + code.encode() + struct.pack(fmt, x)
    if ty == 'f':
        return t + b'f' + struct.pack('<f', float(v))
    if ty == 'A':
        return t + b'A' + v.encode()[:1]
    return t + b'Z' + v.encode() + b'\0'


def encode_record(line, ref_ids):
    f = line.rstrip('\n').split('\t')
    qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
    rid = ref_ids.get(rname, -1)
    nid = rid if rnext == '=' else ref_ids.get(rnext, -1)
    pos0, pnext0 = int(pos) - 1, int(pnext) - 1
    ops = []
    rlen = 0
    if cigar != '*':
        n = 0
        for ch in cigar:
            if ch.isdigit():
                n = n * 10 + ord(ch) - 48
            else:
                ops.append(n << 4 | _CIG[ch])
                if ch in 'MDN=X': rlen += n
                n = 0
    l_seq = 0 if seq == '*' else len(seq)
    end = pos0 + (rlen if rlen else 1)
    name = qname.encode() + b'\0'
    body = struct.pack('<iiBBHHHIiii', rid, pos0, len(name), int(mapq), _reg2bin(max(pos0, 0), max(end, 1)) if pos0 >= 0 else 4680,
                       len(ops), int(flag), l_seq, nid, pnext0, int(tlen)) + name
    body += struct.pack(f'<{len(ops)}I', *ops)
    if l_seq:
        nib = [_SEQ.get(c, 15) for c in seq.upper()]
        if l_seq & 1: nib.append(0)
        body += bytes((nib[i] << 4 | nib[i + 1]) for i in range(0, len(nib), 2))
        body += bytes([0xff] * l_seq) if qual == '*' else bytes((ord(c) - 33) for c in qual)
    for tag in f[11:]:
        body += _aux(tag)
    return struct.pack('<i', len(body)) + body


def sam_to_bam(sam_lines, bam_path, threads=1, level=6):
    """Encode an iterable of SAM text lines (header first) into a BAM file."""
    header, refs = [], []
    pool = ThreadPoolExecutor(max(1, threads))
    pending = []
    with open(bam_path, 'wb') as out:
        buf = bytearray()
        started = False
        ref_ids = {}

        def flush(force=False):
            nonlocal buf
            while len(buf) >= 0xff00 or (force and buf):
                chunk, buf = bytes(buf[:0xff00]), buf[0xff00:]
                pending.append(pool.submit(_bgzf_block, chunk, level))
                if len(pending) > 64:
                    out.write(pending.pop(0).result())

        def start():
            text = ''.join(header).encode()
            h = b'BAM\x01' + struct.pack('<i', len(text)) + text + struct.pack('<i', len(refs))
            for nm, ln in refs:
                h += struct.pack('<i', len(nm) + 1) + nm.encode() + b'\0' + struct.pack('<i', ln)
            buf.extend(h)
        for line in sam_lines:
            if not started and line.startswith('@'):
                header.append(line)
                if line.startswith('@SQ'):
                    d = dict(x.split(':', 1) for x in line.rstrip('\n').split('\t')[1:])
                    ref_ids[d['SN']] = len(refs)
                    refs.append((d['SN'], int(d['LN'])))
                continue
            if not started:
                start(); started = True
            if line.strip():
                buf.extend(encode_record(line, ref_ids))
                flush()
        if not started:
            start()
        flush(force=True)
        for p in pending:
            out.write(p.result())
        out.write(_EOF)
    pool.shutdown()


def sam_stream_to_bam(argv, bam_path, threads, log_fd, index=None, device=0):
    """Run the aligner with its SAM stream piped into the BAM encoder."""
    r, w = os.pipe()
    result = {}

    def produce():
        try:
            result['rc'], result['stats'] = _native.mem_main(argv, index=index, device=device, out_fd=w, log_fd=log_fd)
        finally:
            os.close(w)
    t = threading.Thread(target=produce)
    t.start()
    with os.fdopen(r, 'r') as sam:
        sam_to_bam(sam, bam_path, threads)
    t.join()
    return result.get('rc', 1), result.get('stats', {})
