"""SAM -> BAM for `-O` (the reference pipes `bwa mem` into htslib's stream_bam, External/HTSLIB/stream_bam.c).

Both entry points run in the native library (bsbolt_b200/csrc/host_bam.cpp): the BAM records are what htslib's
sam_parse1 + bam_write1 make of the same SAM lines (uncompressed stream byte-identical to the reference's,
tests/test_bam_output.py), BGZF blocks are deflated by a pool of host threads.

* `-OT n` (output_threads) keeps its meaning when n > 1; the reference's default of 1 would cap the output at one
  deflating core, so 1 means "this process's share of the cores" here (BSB_BAM_THREADS overrides).
* BSB_BAM_LEVEL = zlib level 0..9; default -1 = zlib's default, the level stream_bam writes with.
"""
import os

from bsbolt_b200 import _native


def _threads(output_threads):
    if os.environ.get('BSB_BAM_THREADS'):
        return int(os.environ['BSB_BAM_THREADS'])
    return int(output_threads) if output_threads and int(output_threads) > 1 else 0


def _level():
    return int(os.environ.get('BSB_BAM_LEVEL', '-1'))


def sam_file_to_bam(sam_path, bam_path, threads=1):
    """Encode a SAM file (header first) into a BAM file; returns the number of records."""
    fd = os.open(sam_path, os.O_RDONLY)
    try:
        return _native.stream_bam(fd, bam_path, _threads(threads), _level())
    finally:
        os.close(fd)


def sam_stream_to_bam(argv, bam_path, threads, log_fd, index=None, device=0):
    """Run the aligner with its records written as BAM (no SAM text leaves the library)."""
    if isinstance(index, _native.MultiIndex):
        return _native.mem_main_multi_bam(argv, bam_path, index, threads=_threads(threads), level=_level(), log_fd=log_fd)
    return _native.mem_main_bam(argv, bam_path, index=index, device=device, threads=_threads(threads), level=_level(), log_fd=log_fd)
