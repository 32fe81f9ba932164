#!/usr/bin/env python3
"""Golden BAM streams for the BAM writer (bsbolt_b200/csrc/host_bam.cpp), made by the REFERENCE's own encoder.

Runs oracle/_ref/stream_bam (htslib's sam_read1 -> sam_write1, built from /root/reference/bsbolt/External/HTSLIB by
oracle/Makefile) on every golden SAM of this directory and on bam_edge.sam (hand-written corner cases), inflates the
BGZF output and records length + sha256 of the uncompressed BAM stream in bam_golden.json. Only runs in the build
container (needs oracle/_ref/stream_bam); the tests read the committed JSON.
"""
import glob
import gzip
import hashlib
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
STREAM_BAM = os.path.join(HERE, '..', '..', 'oracle', '_ref', 'stream_bam')


def edge_sam():
    """Records that exercise every branch of sam_parse1 the aligner's output can reach, and a few it cannot."""
    long_seq = 'ACGTN' * 14000          # one record larger than a BGZF block
    H = ['@HD\tVN:1.0\tSO:unsorted', '@SQ\tSN:chr1\tLN:100000', '@SQ\tSN:chr2\tLN:536870912', '@PG\tID:bwa\tPN:bwa\tVN:x\tCL:y z']
    R = [
        'r0\t0\tchr1\t100\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\tNM:i:0\tMD:Z:10\tXB:Z:zz3ZZ2z\tAS:i:10\tXS:i:0\tYS:Z:W_C2T\tXG:Z:CT',
        'r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGTN\t!!~~I\tAS:i:0\tYS:Z:WC',
        'r2\t77\t*\t0\t0\t*\t*\t0\t0\tacgtnRYK\t*\tAS:i:0',
        'r2\t141\t*\t0\t0\t*\t*\t0\t0\t*\t*',
        'r3\t73\tchr1\t5000\t0\t3S4M2I1D3M1H\t=\t5000\t0\tACGTACGTACGT\tABCDEFGHIJKL\tNM:i:255\tXI:i:256\tXJ:i:65535\tXK:i:65536\tXL:i:4294967295',
        'r3\t133\tchr1\t5000\t0\t*\t=\t5000\t0\tACG\tIII\tXa:i:-1\tXb:i:-128\tXc:i:-129\tXd:i:-32768\tXe:i:-32769\tXf:f:0.25\tXg:A:Q\tXh:Z:\tXi:H:1AE301',
        'r4\t99\tchr2\t536870000\t60\t5M\tchr1\t1\t-536869999\tACGTA\tIIIII\tXA:Z:chr1,-3576,5M,1;chr2,+17,5M,0;',
        'r5\t16\tchrUnknown\t17\t3\t4M\t*\t0\t0\tAAAA\tIIII',
        'r6\t0\tchr1\t0\t3\t4M\tchr1\t0\t0\tAAAA\tIIII',
        'r7\t0\tchr1\t16384\t3\t1M\t=\t16385\t2\tA\tI',
        'r8\t0\tchr1\t16380\t3\t10M\t=\t1\t0\tAAAAAAAAAA\tIIIIIIIIII',
        'r9\t0\tchr1\t131070\t3\t2M3N2M\t=\t1\t0\tAAAA\tIIII',
        'r10\t256\tchr1\t7\t0\t2H3=1X\t*\t0\t0\tACGT\tIIII\tSA:Z:chr1,9,+,4M,0,0;',
        'rl\t0\tchr1\t1\t60\t70000M\t*\t0\t0\t' + long_seq + '\t' + 'I' * len(long_seq) + '\tNM:i:0',
        'r11\t0\tchr1\t9\t60\t3M\t*\t0\t0\tACG\tIII\tRG:Z:grp\tXR:Z:comment with spaces',
    ]
    return '\n'.join(H + R) + '\n'


def raw_bam_of(sam_bytes):
    with tempfile.TemporaryDirectory() as d:
        bam = os.path.join(d, 'x.bam')
        subprocess.run([STREAM_BAM, '-o', bam], input=sam_bytes, check=True, stderr=subprocess.DEVNULL)
        return gzip.open(bam, 'rb').read()


def main():
    open(os.path.join(HERE, 'bam_edge.sam'), 'w').write(edge_sam())
    out = {}
    for g in sorted(glob.glob(os.path.join(HERE, '*.sam.gz'))) + [os.path.join(HERE, 'bam_edge.sam')]:
        sam = gzip.open(g, 'rb').read() if g.endswith('.gz') else open(g, 'rb').read()
        raw = raw_bam_of(sam)
        out[os.path.basename(g)] = {'raw_len': len(raw), 'raw_sha256': hashlib.sha256(raw).hexdigest()}
        print(os.path.basename(g), out[os.path.basename(g)])
    json.dump({'made_by': 'oracle/_ref/stream_bam (reference htslib, HTSLIB/stream_bam.c)', 'streams': out},
              open(os.path.join(HERE, 'bam_golden.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
