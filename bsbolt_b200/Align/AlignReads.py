"""Drop-in for bsbolt/Align/AlignReads.py:18-87 (class BisulfiteAlignmentAndProcessing).

Same constructor, same `align_reads()`, same `mapping_statistics` keys and the same two exceptions.
Where the reference spawns `bwa mem ... | stream_bam`, this class hands the identical argv to the
in-process GPU aligner (bsb_mem_main in include/bsbolt_b200.h); SAM text and `BSStat` log lines keep
their formats, so anything that consumed the reference's streams keeps working.
"""
import os
import sys
import tempfile
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
        if isinstance(self.device, (list, tuple)) and len(self.device) > 1:
            return self._align_multi_gpu(argv)
        if isinstance(self.device, (list, tuple)):
            self.device = self.device[0]
        with tempfile.TemporaryFile(mode='w+') as log:
            if self.output_to_stdout:
                rc, stats = _native.mem_main(argv, index=self.index, device=self.device, out_fd=1, log_fd=log.fileno())
            else:
                from bsbolt_b200.Utils.BamOutput import sam_stream_to_bam
                rc, stats = sam_stream_to_bam(argv, f'{self.output}.bam', self.output_threads, log.fileno(),
                                              index=self.index, device=self.device)
            log.seek(0)
            for alignment_info in log:
                if alignment_info[0:7] == 'BSStat ':
                    category, count = alignment_info.replace('BSStat ', '').split(': ')
                    self.mapping_statistics[category] += int(count)
                    print(alignment_info.replace('BSStat ', '').strip(), file=sys.stderr)
                else:
                    print(alignment_info.strip(), file=sys.stderr)
        self.run_statistics = stats
        if rc:
            print(rc, file=sys.stderr)
            raise BisulfiteAlignmentError(_native.last_error())

    def _consume_log(self, lines):
        for alignment_info in lines:
            if alignment_info[0:7] == 'BSStat ':
                category, count = alignment_info.replace('BSStat ', '').split(': ')
                self.mapping_statistics[category] += int(count)
                print(alignment_info.replace('BSStat ', '').strip(), file=sys.stderr)
            else:
                print(alignment_info.strip(), file=sys.stderr)

    def _align_multi_gpu(self, argv):
        """One worker process per GPU; batch b is aligned on GPU b mod G; parts merged in input order."""
        import subprocess
        from bsbolt_b200.shard import merge_shards
        devices = list(self.device)
        with tempfile.TemporaryDirectory() as d:
            procs = []
            for i, dev in enumerate(devices):
                cmd = [sys.executable, '-m', 'bsbolt_b200._shard_worker', str(dev), str(i), str(len(devices)),
                       f'{d}/p{i}.sam', f'{d}/p{i}.idx', f'{d}/p{i}.log', '--'] + argv
                procs.append(subprocess.Popen(cmd))
            rcs = [p.wait() for p in procs]
            for i in range(len(devices)):
                self._consume_log(open(f'{d}/p{i}.log'))
            if any(rcs):
                raise BisulfiteAlignmentError(f'shard worker failed: {rcs}')
            sams, parts = [f'{d}/p{i}.sam' for i in range(len(devices))], [f'{d}/p{i}.idx' for i in range(len(devices))]
            if self.output_to_stdout:
                merge_shards(sams, parts, sys.stdout.buffer)
                sys.stdout.buffer.flush()
            else:
                from bsbolt_b200.Utils.BamOutput import sam_file_to_bam
                with open(f'{d}/merged.sam', 'wb') as o:
                    merge_shards(sams, parts, o)
                sam_file_to_bam(f'{d}/merged.sam', f'{self.output}.bam', self.output_threads)
