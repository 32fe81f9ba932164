"""`bsbolt Index` on the GPU: whole-genome, bed-masked whole-genome and RRBS (in-silico digested) databases.

Reference: bsbolt/Index/WholeGenomeIndex.py:34-124 (whole genome, optional `-MR` bed of mappable regions),
bsbolt/Index/RRBSIndex.py:46-169 + RRBSCutSites.py (restriction fragments inside a size window stay, everything
else becomes '-'), bsbolt/Index/IndexOutput.py:38-74 (`<DB>/BSB_ref.fa`, one line per contig, per-contig pickles,
`mappable_regions.bed.gz`, then `bwa index -a bwtsw`). Here the same `BSB_ref.fa` text is written and the index
files come from the GPU builder (bsb_index_build), byte-identical to the reference's: a masked base is just
another non-ACGT character to the packer (bntseq.c:261-299: random 2-bit code from lrand48, one `.amb` hole per
run). The pickles that only the downstream methylation caller reads are written too, so a database made here
is a complete drop-in.
"""
import gzip
import os
import pickle
import re

import numpy as np

from bsbolt_b200 import _native

_IUPAC = {'R': 'AG', 'Y': 'CT', 'S': 'GC', 'W': 'AT', 'K': 'GT', 'M': 'AC', 'B': 'CGT', 'D': 'AGT', 'H': 'ACT', 'V': 'ACG',
          'N': 'ACGT'}
_COMP = str.maketrans('ATGC', 'TACG')


def _contigs(reference_fasta):
    """(contig id, list of sequence lines) as OpenFasta + the builders' loops see them (Utils/FastaIterator.py:21-38):
    a line holding '>' anywhere is a label, only the newline is stripped."""
    op = gzip.open if str(reference_fasta).endswith('.gz') else open
    cid, parts = None, []
    with op(reference_fasta, 'rt') as f:
        for line in f:
            line = line.replace('\n', '')
            if '>' in line:
                if cid:
                    yield cid, parts
                cid, parts = line.replace('>', '').split()[0], []
            else:
                parts.append(line)
    yield cid, parts


def restriction_sites(cut_format='C-CGG'):
    """recognition sequence -> cut offset, in the reference's insertion order (RRBSCutSites.py:22-83)"""
    sites = {}
    for site in cut_format.upper().replace(' ', '').split(','):
        fwd_off = rev_off = 0
        if '-' in site:
            fwd_off = site.index('-')
            rev_off = site[::-1].translate(_COMP).index('-')
        fwd = ['']
        for nt in site.replace('-', ''):
            fwd = [s + x for x in _IUPAC.get(nt, nt) for s in fwd]
        for f in fwd:
            r = f[::-1].translate(_COMP)
            sites[f] = fwd_off
            if f != r:
                sites[r] = rev_off
    return sites


def rrbs_regions(seq, sites, lower_bound=30, upper_bound=500):
    """mappable (start, end) pairs of one contig (RRBSIndex.py:103-139): consecutive restriction sites whose
    fragment length lies inside the size window, widened by two bases on both sides"""
    hits = []
    for site, off in sites.items():
        hits.extend((m.start(), off, len(site)) for m in re.finditer(site, seq))
    hits.sort(key=lambda h: h[0])
    out = []
    for (p0, off0, _), (p1, off1, l1) in zip(hits, hits[1:]):
        start = p0 + off0
        end = p1 + (off1 if off1 else l1)
        if lower_bound <= end - start <= upper_bound:
            out.append((start - 2, end + 2))
    return out


def mask_outside(seq, regions):
    """`seq` with everything outside `regions` replaced by '-'. Reproduces the reference's position loop
    (RRBSIndex.py:141-169, WholeGenomeIndex.py:94-121) region by region: the loop holds ONE current region, moves
    on at most once per position (the first position beyond the current end) and never past the last one; a position
    is kept when start <= position <= end of the region current at that position."""
    n = len(seq)
    src = np.frombuffer(seq.encode('latin-1'), dtype=np.uint8)
    out = np.full(n, ord('-'), dtype=np.uint8)
    a = 0                                   # first position at which region k is the current one
    for k, (start, end) in enumerate(regions):
        last = k == len(regions) - 1
        if last:
            b = n
        elif k == 0:
            b = max(0, end + 1)             # position 0 may already lie beyond the first region
        else:
            b = max(a + 1, end + 1)
        lo, hi = max(a, start, 0), min(b - 1, end, n - 1)
        if hi >= lo:
            out[lo:hi + 1] = src[lo:hi + 1]
        a = b
        if a >= n:
            break
    return out.tobytes().decode('latin-1')


def read_mappable_bed(bed_file):
    """`-MR` regions per contig, widened by two bases and sorted by start (WholeGenomeIndex.py:66-92)"""
    regions = {}
    with open(bed_file) as f:
        for line in f:
            if line:
                chrom, start, end = line.replace('\n', '').split('\t')[0:3]
                regions.setdefault(chrom, []).append((int(start) - 2, int(end) + 2))
    for v in regions.values():
        v.sort(key=lambda r: r[0])
    return regions


def _open_db(genome_database):
    if not genome_database.endswith('/'):
        genome_database += '/'
    os.makedirs(genome_database, exist_ok=True)
    return genome_database, f'{genome_database}BSB_ref.fa'


def _dump(genome_database, name, obj):
    with open(f'{genome_database}{name}.pkl', 'wb') as f:
        pickle.dump(obj, f)


def write_database_fasta(reference_fasta, genome_database, mappable_regions=None, ignore_alt=False):
    """the FASTA + pickles of a whole-genome database (optionally bed-masked); no index yet"""
    genome_database, ref = _open_db(genome_database)
    regions = read_mappable_bed(mappable_regions) if mappable_regions else None
    sizes = {}
    with open(ref, 'w') as out:
        for cid, parts in _contigs(reference_fasta):
            if cid is None or (ignore_alt and 'alt' in cid.lower()):
                continue
            seq = ''.join(parts)
            if regions is not None:
                seq = mask_outside(seq, regions[cid]) if cid in regions else '-' * len(seq)
            _dump(genome_database, cid, seq)
            out.write(f'>{cid}\n{seq}\n')
            sizes[cid] = len(seq)
    _dump(genome_database, 'genome_index', sizes)
    return ref


def write_rrbs_database_fasta(reference_fasta, genome_database, lower_bound=40, upper_bound=500, cut_format='C-CGG', ignore_alt=False):
    """the FASTA, pickles and mappable_regions.bed.gz of an RRBS database (RRBSIndex.py:46-101); no index yet"""
    genome_database, ref = _open_db(genome_database)
    sites = restriction_sites(cut_format)
    sizes, bed = {}, []
    with open(ref, 'w') as out:
        for cid, parts in _contigs(reference_fasta):
            if cid is None or (ignore_alt and 'alt' in cid.lower()):
                continue
            seq = ''.join(p.upper() for p in parts)
            sizes[cid] = len(seq)
            _dump(genome_database, cid, seq)              # the unmasked sequence, like the reference
            regions = rrbs_regions(seq, sites, lower_bound, upper_bound) or [(1, 80)]
            bed.extend(f'{cid}\t{s}\t{e}\n' for s, e in regions)
            out.write(f'>{cid}\n{mask_outside(seq, regions)}\n')
    with gzip.open(f'{genome_database}mappable_regions.bed.gz', 'wb') as f:
        for line in bed:
            f.write(line.encode('UTF-8'))
    _dump(genome_database, 'genome_index', sizes)
    return ref


def build_database(reference_fasta, genome_database, device=0, ignore_alt=False, mappable_regions=None):
    """`bsbolt Index -G <fasta> -DB <dir> [-MR bed] [-IA]`"""
    ref = write_database_fasta(reference_fasta, genome_database, mappable_regions=mappable_regions, ignore_alt=ignore_alt)
    ms = _native.index_build(ref, ref, device)
    return ref, ms


def build_rrbs_database(reference_fasta, genome_database, device=0, lower_bound=40, upper_bound=500, cut_format='C-CGG', ignore_alt=False):
    """`bsbolt Index -G <fasta> -DB <dir> -rrbs [-rrbs-cut-format C-CGG] [-rrbs-lower 40] [-rrbs-upper 500]`
    (defaults of the command line, bsbolt/Utils/Parser.py:131-140; the RRBSBuild class itself defaults to 30)"""
    ref = write_rrbs_database_fasta(reference_fasta, genome_database, lower_bound, upper_bound, cut_format, ignore_alt)
    ms = _native.index_build(ref, ref, device)
    return ref, ms
