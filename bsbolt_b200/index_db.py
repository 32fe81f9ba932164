"""`bsbolt Index` for whole-genome (unmasked) databases on the GPU.

Reference: bsbolt/Index/WholeGenomeIndex.py:34-53 + bsbolt/Index/IndexOutput.py:38-56 write `<DB>/BSB_ref.fa`
(one line per contig) and run `bwa index -a bwtsw` on it. Here the same FASTA is written and the index files
come from the GPU builder (bsb_index_build), byte-identical to the reference's. The per-contig pickles that only
the downstream methylation caller reads are also written so that a database made here is a complete drop-in.
"""
import gzip
import os
import pickle

from bsbolt_b200 import _native


def build_database(reference_fasta, genome_database, device=0, ignore_alt=False):
    if not genome_database.endswith('/'):
        genome_database += '/'
    os.makedirs(genome_database, exist_ok=True)
    ref = f'{genome_database}BSB_ref.fa'
    op = gzip.open if str(reference_fasta).endswith('.gz') else open
    sizes = {}

    def flush(out, cid, parts):
        if cid is None or (ignore_alt and 'alt' in cid.lower()):
            return
        seq = ''.join(parts)
        out.write(f'>{cid}\n{seq}\n')
        with open(f'{genome_database}{cid}.pkl', 'wb') as f:
            pickle.dump(seq, f)
        sizes[cid] = len(seq)
    with op(reference_fasta, 'rt') as f, open(ref, 'w') as out:
        cid, parts = None, []
        for line in f:
            if line.startswith('>'):
                flush(out, cid, parts)
                cid, parts = line[1:].split()[0], []
            else:
                parts.append(line.strip())
        flush(out, cid, parts)
    with open(f'{genome_database}genome_index.pkl', 'wb') as f:
        pickle.dump(sizes, f)
    ms = _native.index_build(ref, ref, device)
    return ref, ms
