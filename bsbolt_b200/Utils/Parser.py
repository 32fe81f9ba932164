"""`bsbolt Align` and `bsbolt Index` command lines, flag for flag (reference: bsbolt/Utils/Parser.py:31-115, 118-140).

Align is the GPU hot path; Index produces its input (the suffix array is built on the GPU, the files are byte-identical to
`bwa index`). The other bsbolt modules (Simulate, CallMethylation, ...) keep running from the reference package.
"""
import argparse

_strip = lambda x: x.strip()  # noqa: E731

# (flag, kwargs) in the reference's order; defaults are the reference's defaults
ALIGN_FLAGS = [
    ('-F1', dict(type=str, default=None, required=True, help='path to fastq 1')),
    ('-F2', dict(type=str, default=None, help='path to fastq 2')),
    ('-UN', dict(action='store_true', default=False,
                 help='library undirectional, ie. consider PCR products of bisulfite converted DNA')),
    ('-O', dict(type=str, default=None, help='output Prefix')),
    ('-OS', dict(action='store_true', default=False, help='Output alignment to stdout')),
    ('-DB', dict(type=str, default=None, required=True, help='path to bsbolt database')),
    ('-CP', dict(type=float, default=0.5, help='CH conversion proportion threshold [0.5]')),
    ('-CT', dict(type=int, default=5, help='number of CH sites needed to assess read conversion')),
    ('-SP', dict(type=float, default=0.1,
                 help='substitution threshold for read bisulfite conversion patterns (ie C2T, G2A) [0.1]')),
    ('-t', dict(type=int, default=1, help='number of bwa threads [1] (sets the batch size: 10 Mbp x t)')),
    ('-k', dict(type=int, default=19, help='minimum seed length [19]')),
    ('-w', dict(type=int, default=100, help='band width for banded alignment [100]')),
    ('-d', dict(type=int, default=100, help='off-diagonal X-dropoff [100]')),
    ('-r', dict(type=float, default=1.5, help='look for internal seeds inside a seed longer than {-k} * FLOAT [1.5]')),
    ('-y', dict(type=int, default=20, help='seed occurrence for the 3rd round seeding [20]')),
    ('-c', dict(type=int, default=500, help='skip seeds with more than INT occurrences [500]')),
    ('-D', dict(type=float, default=0.50,
                help='drop chains shorter than FLOAT fraction of the longest overlapping chain [0.50]')),
    ('-W', dict(type=int, default=0, help='discard a chain if seeded bases shorter than INT [0]')),
    ('-m', dict(type=int, default=50, help='perform at most INT rounds of mate rescues for each read [50]')),
    ('-S', dict(action='store_true', default=False, help='skip mate rescue')),
    ('-P', dict(action='store_true', default=False, help='skip pairing; mate rescue performed unless -S also in use')),
    ('-A', dict(type=int, default=1, help='score for a sequence match, which scales options -TdBOELU unless overridden [1]')),
    ('-B', dict(type=int, default=4, help='penalty for a mismatch [4]')),
    ('-INDEL', dict(type=_strip, default='6,6', help='gap open penalties for deletions and insertions [6,6]')),
    ('-E', dict(type=_strip, default='1,1', help="gap extension penalty; a gap of size k cost '{-O} + {-E}*k' [1,1]")),
    ('-L', dict(type=_strip, default='30,30', help="penalty for 5'- and 3'-end clipping [30,30]")),
    ('-U', dict(type=int, default='17', help='penalty for an unpaired read pair [17]')),
    ('-p', dict(action='store_true', default=False, help='smart pairing (ignoring in2.fq)')),
    ('-R', dict(type=str, default=None, help="read group header line such as '@RG\\tID:foo\\tSM:bar' [null]")),
    ('-H', dict(type=str, default=None, help='insert STR to header if it starts with @; or insert lines in FILE [null]')),
    ('-j', dict(action='store_true', default=False,
                help='treat ALT contigs as part of the primary assembly (i.e. ignore <idxbase>.alt file)')),
    ('-T', dict(type=int, default=10, help='minimum score to output [10], set based on read length')),
    ('-XA', dict(type=_strip, default='100,200',
                 help='if there are <INT hits with score >80 percent of the max score, output all in XA [100,200]')),
    ('-DR', dict(type=float, default=0.95, help='drop ratio for alternative hits reported in XA tag [0.95]')),
    ('-M', dict(action='store_true', default=False, help='mark shorter split hits as secondary')),
    ('-I', dict(type=_strip, default=None,
                help='mean, standard deviation, max and min of the insert size distribution. FR orientation only. [inferred]')),
    ('-OT', dict(type=int, default=1, help='number of threads of bam output threads[1]')),
]

# extensions of this build (not in the reference parser)
EXTRA_FLAGS = [
    ('-K', dict(type=int, default=None, help='process INT input bases in each batch regardless of -t (bwa mem -K; reproducible batches)')),
    ('-GPU', dict(type=lambda x: [int(v) for v in str(x).split(',')], default=[0],
                  help='CUDA device(s) to align on, comma separated; with several devices batch b runs on device b mod G [0]')),
]

parser = argparse.ArgumentParser(description='bsbolt_b200: GPU drop-in for the bsbolt Align (and Index) modules', prog='python -m bsbolt_b200')
subparsers = parser.add_subparsers(description='module', metavar='Align | Index', dest='subparser_name')
align_parser = subparsers.add_parser('Align', help='Alignment', add_help=True)
for _flag, _kw in ALIGN_FLAGS + EXTRA_FLAGS:
    align_parser.add_argument(_flag, **_kw)

# `bsbolt Index` (reference flags and defaults); -B is accepted and ignored: it sizes the blocks of the CPU bwtsw algorithm
INDEX_FLAGS = [
    ('-G', dict(type=str, required=True, help='Path to reference genome fasta file, fasta file should contain all contigs')),
    ('-DB', dict(type=str, required=True, help='Path to index directory, will create directory if folder does not exist')),
    ('-B', dict(type=int, default=10000000, help='Block size for the bwtsw algorithm of the reference (ignored: the suffix array is built on the GPU)')),
    ('-MR', dict(type=str, default=None, help='Path to bed file of mappable regions. Index will be built using masked contig sequence')),
    ('-IA', dict(action='store_true', default=False, help='ignore alt contigs during index construction')),
    ('-rrbs', dict(action='store_true', default=False, help='Generate a Reduced Representative Bisulfite Sequencing (RRBS) index')),
    ('-rrbs-cut-format', dict(default='C-CGG', help='Cut format for the RRBS database, default= C-CGG (MSPI); several enzymes comma separated')),
    ('-rrbs-lower', dict(type=int, default=40, help='Lower bound fragment size to consider for the RRBS index, default = 40')),
    ('-rrbs-upper', dict(type=int, default=500, help='Upper bound fragment size to consider for the RRBS index, default = 500')),
]
index_parser = subparsers.add_parser('Index', help='Index Generation', add_help=True)
for _flag, _kw in INDEX_FLAGS:
    index_parser.add_argument(_flag, **_kw)
index_parser.add_argument('-GPU', type=int, default=0, help='CUDA device that builds the suffix array [0]')
