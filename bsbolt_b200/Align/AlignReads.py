"""Drop-in for bsbolt/Align/AlignReads.py:18-87 (class BisulfiteAlignmentAndProcessing).

Same constructor, same `align_reads()`, same `mapping_statistics` keys and the same two exceptions.
Where the reference spawns `bwa mem ... | stream_bam`, this class hands the identical argv to the
in-process GPU aligner (bsb_mem_main in include/bsbolt_b200.h); SAM text and `BSStat` log lines keep
their formats, so anything that consumed the reference's streams keeps working.
"""
import os
import sys
import threading
from typing import List

from bsbolt_b200 import _native


class BisulfiteAlignmentError(Exception):
    """Error in alignment"""
    pass


class AlignmentCompressionError(Exception):
    """Error in read compression"""
    pass


class BisulfiteAlignmentAndProcessing:
    """Read alignment and processing on the GPU.

    Params (identical to the reference):

    * *alignment_commands (list)*: bwa alignment commands ([bwa_path, 'mem', ...]; element 0 is ignored)
    * *output (str)*: output prefix
    * *output_threads (int)*: number of threads available for bam output
    * *output_to_stdout (bool)*: output alignments to stdout

    Attributes:

    * *self.mapping_statistics (dict)*: alignment run statistics
    * *self.run_statistics (dict)*: device timings of the run (extension)
    """

    def __init__(self, alignment_commands: List[str], output: str = None, output_threads: int = 1,
                 output_to_stdout: bool = False, device: int = 0, index=None):
        self.alignment_commands = alignment_commands
        self.output = output
        self.output_threads = output_threads
        self.output_to_stdout = output_to_stdout
        self.device = device
        self.index = index
        self.mapping_statistics = dict(TotalReads=0, TotalAlignments=0, BSAmbiguous=0, C_C2T=0, C_G2A=0,
                                       W_C2T=0, W_G2A=0, Unaligned=0)
        self.run_statistics = {}

    def align_reads(self):
        """Launch the alignment. SAM goes to stdout (-OS) or is compressed to <output>.bam"""
        if self.output is not None and '/' in self.output:
            assert os.path.exists('/'.join(self.output.split('/')[0:-1])), f"output path {self.output} not valid"
        argv = [str(a) for a in self.alignment_commands[1:]]
        if not argv or argv[0] != 'mem':
            raise BisulfiteAlignmentError('alignment_commands must be [bwa, "mem", ...]')
        sys.stdout.flush()
        devices = list(self.device) if isinstance(self.device, (list, tuple)) else [self.device]
        multi = None
        if len(devices) > 1:
            # several GPUs, ONE process: the input is read and cut into batches once, batch b runs on device b mod G against
            # that device's resident copy of the index, records leave in input order (bsb_mem_main_multi); no temp files
            multi = self.index if isinstance(self.index, _native.MultiIndex) else _native.MultiIndex(self._idxbase(argv), devices)
        # the aligner's log is consumed line by line while it runs, like the reference reads bwa's stderr
        # (AlignReads.py:61-78): progress and per-batch BSStat counters appear as the batches finish
        rd, wr = os.pipe()
        reader = threading.Thread(target=self._consume_log, args=(os.fdopen(rd, 'r', errors='replace'),), daemon=True)
        reader.start()
        try:
            if self.output_to_stdout:
                if multi is not None:
                    rc, stats = _native.mem_main_multi(argv, multi, out_fd=1, log_fd=wr)
                else:
                    rc, stats = _native.mem_main(argv, index=self.index, device=devices[0], out_fd=1, log_fd=wr)
            else:
                from bsbolt_b200.Utils.BamOutput import sam_stream_to_bam
                rc, stats = sam_stream_to_bam(argv, f'{self.output}.bam', self.output_threads, wr,
                                              index=multi if multi is not None else self.index, device=devices[0])
        finally:
            os.close(wr)
            reader.join()
            if multi is not None and multi is not self.index:
                multi.close()
        self.run_statistics = stats
        if rc:
            print(rc, file=sys.stderr)
            raise BisulfiteAlignmentError(_native.last_error())

    @staticmethod
    def _idxbase(argv):
        """the index prefix in a `bwa mem` argv: the first positional argument (options as in fastmap.c:113-190)"""
        flags = set('51qpaMCSPVYjuz')   # the options of `bwa mem` that take no value (getopt string of main_mem)
        i = 1
        while i < len(argv):
            a = argv[i]
            if a.startswith('-') and len(a) > 1:
                k = 1
                while k < len(a) and a[k] in flags:   # bundled flags, e.g. -MY
                    k += 1
                i += 1 if (k == len(a) or k + 1 < len(a)) else 2   # a value glued to its option (-k19) or the next word
            else:
                return a
        raise BisulfiteAlignmentError('no index in the alignment command')

    def _consume_log(self, lines):
        with lines:
            for alignment_info in lines:
                if alignment_info[0:7] == 'BSStat ':
                    category, count = alignment_info.replace('BSStat ', '').split(': ')
                    self.mapping_statistics[category] += int(count)
                    print(alignment_info.replace('BSStat ', '').strip(), file=sys.stderr)
                else:
                    print(alignment_info.strip(), file=sys.stderr)
