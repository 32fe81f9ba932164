"""ctypes binding of libbsbolt_b200.so (C ABI in include/bsbolt_b200.h). No torch, no fallback."""
import ctypes as C
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libbsbolt_b200.so')
_lib = None


class NativeLibraryMissing(RuntimeError):
    """libbsbolt_b200.so has not been built (run `make -C bsbolt_b200/csrc` or __graft_entry__.build())."""


class RunStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('total_reads', 'total_alignments', 'w_c2t', 'w_g2a', 'c_c2t', 'c_g2a',
                                         'unaligned', 'bs_ambiguous', 'n_batches', 'n_entries')] + \
               [('sec_total', C.c_double), ('sec_align', C.c_double),
                ('ms_h2d', C.c_double), ('ms_kernels', C.c_double), ('ms_d2h', C.c_double),
                ('ms_stage', C.c_double * 8)] + \
               [(n, C.c_int64) for n in ('n_seeds', 'h2d_bytes', 'd2h_bytes', 'kernel_launches')] + \
               [(n, C.c_double) for n in ('sec_read', 'sec_format', 'sec_write', 'ms_select', 'ms_tasks')] + [('n_tasks', C.c_int64), ('sec_resident', C.c_double)] + \
               [(n, C.c_int64) for n in ('fm_extensions', 'fm_two_block', 'fm_block_bytes', 'dp_cells_extend', 'fm_two_block_ref')] + \
               [('sec_plan', C.c_double), ('sec_fill', C.c_double), ('rescue_pairs', C.c_int64), ('rescue_jobs', C.c_int64)] + \
               [('ms_text', C.c_double), ('ms_bam', C.c_double), ('bam_raw_bytes', C.c_int64), ('bam_bgzf_bytes', C.c_int64), ('bam_blocks', C.c_int64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == 'ms_stage' else v
        return d


class Read(C.Structure):
    _fields_ = [('name', C.c_char_p), ('comment', C.c_char_p), ('seq', C.c_char_p), ('qual', C.c_char_p)]


EXPORTS = ('bsb_version', 'bsb_last_error', 'bsb_device_count', 'bsb_run_stats_size', 'bsb_index_load', 'bsb_index_free',
           'bsb_index_hbm_bytes', 'bsb_index_n_contigs', 'bsb_mem_main', 'bsb_batch_create', 'bsb_batch_align',
           'bsb_batch_sam', 'bsb_batch_n_entries', 'bsb_batch_free', 'bsb_sam_header', 'bsb_index_build',
           'bsb_mem_main_bam', 'bsb_stream_bam', 'bsb_index_clone', 'bsb_mem_main_multi', 'bsb_mem_main_multi_bam', 'bsb_random_sector_peak')


def lib():
    """Load the shared library once; fail loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeLibraryMissing(f'{_LIB_PATH} not found: build it with `make -C bsbolt_b200/csrc`; '
                                   'this aligner has no CPU fallback')
    L = C.CDLL(_LIB_PATH)
    L.bsb_version.restype = C.c_char_p
    L.bsb_last_error.restype = C.c_char_p
    L.bsb_device_count.restype = C.c_int
    L.bsb_run_stats_size.restype = C.c_size_t
    if L.bsb_run_stats_size() != C.sizeof(RunStats):   # a stale library next to a newer binding (or the reverse) would scribble over memory
        raise NativeLibraryMissing(f'{_LIB_PATH} was built from another include/bsbolt_b200.h (bsb_run_stats_t is '
                                   f'{L.bsb_run_stats_size()} bytes there, {C.sizeof(RunStats)} here): rebuild with `make -C bsbolt_b200/csrc`')
    L.bsb_index_load.restype = C.c_void_p
    L.bsb_index_load.argtypes = [C.c_char_p, C.c_int]
    L.bsb_index_free.argtypes = [C.c_void_p]
    L.bsb_index_clone.restype = C.c_void_p
    L.bsb_index_clone.argtypes = [C.c_void_p, C.c_int]
    L.bsb_mem_main_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(RunStats)]
    L.bsb_mem_main_multi_bam.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(RunStats)]
    L.bsb_index_hbm_bytes.restype = C.c_int64
    L.bsb_index_hbm_bytes.argtypes = [C.c_void_p]
    L.bsb_index_n_contigs.argtypes = [C.c_void_p]
    L.bsb_mem_main.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(RunStats)]
    L.bsb_mem_main_bam.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(RunStats)]
    L.bsb_stream_bam.restype = C.c_int64
    L.bsb_stream_bam.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int]
    L.bsb_batch_create.restype = C.c_void_p
    L.bsb_batch_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(Read), C.POINTER(Read)]
    L.bsb_batch_align.argtypes = [C.c_void_p, C.c_int64, C.POINTER(RunStats)]
    L.bsb_batch_sam.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.POINTER(RunStats)]
    L.bsb_batch_n_entries.argtypes = [C.c_void_p]
    L.bsb_batch_free.argtypes = [C.c_void_p]
    L.bsb_sam_header.restype = C.c_char_p
    L.bsb_sam_header.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
    L.bsb_index_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    L.bsb_random_sector_peak.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = L
    return L


def last_error():
    return lib().bsb_last_error().decode(errors='replace')


def _argv(args):
    arr = (C.c_char_p * len(args))(*[a.encode() if isinstance(a, str) else a for a in args])
    return arr


class Index:
    """Index of `bsbolt Index` resident in one GPU's HBM."""

    def __init__(self, idxbase, device=0, _handle=None):
        self._h = _handle if _handle is not None else lib().bsb_index_load(str(idxbase).encode(), int(device))
        if not self._h:
            raise RuntimeError(last_error())
        self.device = device

    def clone(self, device):
        """One more resident copy of this index, on another device (host-side tables shared)."""
        h = lib().bsb_index_clone(self._h, int(device))
        if not h:
            raise RuntimeError(last_error())
        return Index(None, device, _handle=h)

    @property
    def hbm_bytes(self):
        return lib().bsb_index_hbm_bytes(self._h)

    def close(self):
        if self._h:
            lib().bsb_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def index_build(fasta, prefix, device=0):
    """GPU replacement of `bwa index -a bwtsw` (writes <prefix>.pac .opac .ann .amb .bwt .sa). Returns device ms."""
    ms = C.c_double()
    if lib().bsb_index_build(str(fasta).encode(), str(prefix).encode(), int(device), C.byref(ms)):
        raise RuntimeError(last_error())
    return ms.value


def random_sector_peak(device=0, footprint_bytes=1 << 29):
    """GB/s of random 32-byte sector reads over about footprint_bytes: (independent loads, one dependent load per thread)."""
    a, b = C.c_double(), C.c_double()
    if lib().bsb_random_sector_peak(int(device), int(footprint_bytes), C.byref(a), C.byref(b)):
        raise RuntimeError(last_error())
    return a.value, b.value


def mem_main(argv, index=None, device=0, out_fd=1, log_fd=2):
    """`bwa mem` as BSBolt runs it. argv[0] must be 'mem'. Returns (return code, stats dict)."""
    st = RunStats()
    arr = _argv(argv)
    rc = lib().bsb_mem_main(index._h if index is not None else None, int(device), len(argv), arr, int(out_fd), int(log_fd), C.byref(st))
    return rc, st.as_dict()


def mem_main_bam(argv, bam_path, index=None, device=0, threads=0, level=-1, log_fd=2):
    """`bwa mem ... | stream_bam -@ threads -o bam_path` in one call: records leave as BAM. threads <= 0: this process's
    share of the cores; level: zlib level, -1 = default (as stream_bam). Returns (return code, stats dict)."""
    st = RunStats()
    arr = _argv(argv)
    rc = lib().bsb_mem_main_bam(index._h if index is not None else None, int(device), len(argv), arr, str(bam_path).encode(),
                                int(threads), int(level), int(log_fd), C.byref(st))
    return rc, st.as_dict()


class MultiIndex:
    """The same index resident on several GPUs of one box: reads shard by batch, the index is replicated, no collective."""

    def __init__(self, idxbase, devices):
        devices = [int(d) for d in devices]
        if not devices:
            raise ValueError('MultiIndex needs at least one device')
        self.parts = [Index(idxbase, devices[0])]
        for d in devices[1:]:
            self.parts.append(self.parts[0].clone(d))
        self.devices = devices

    @property
    def hbm_bytes(self):
        return self.parts[0].hbm_bytes

    def handles(self):
        return (C.c_void_p * len(self.parts))(*[p._h for p in self.parts])

    def close(self):
        for p in reversed(self.parts):
            p.close()
        self.parts = []


def mem_main_multi(argv, multi_index, out_fd=1, log_fd=2):
    """`bwa mem` over several GPUs: one reader, batch b on device b mod G, output in input order (bsb_mem_main_multi)."""
    st = RunStats()
    rc = lib().bsb_mem_main_multi(multi_index.handles(), len(multi_index.parts), len(argv), _argv(argv), int(out_fd), int(log_fd), C.byref(st))
    return rc, st.as_dict()


def mem_main_multi_bam(argv, bam_path, multi_index, threads=0, level=-1, log_fd=2):
    st = RunStats()
    rc = lib().bsb_mem_main_multi_bam(multi_index.handles(), len(multi_index.parts), len(argv), _argv(argv), str(bam_path).encode(),
                                      int(threads), int(level), int(log_fd), C.byref(st))
    return rc, st.as_dict()


def stream_bam(in_fd, bam_path, threads=0, level=-1):
    """SAM text on a file descriptor -> BAM file (host only). Returns the number of records."""
    n = lib().bsb_stream_bam(int(in_fd), str(bam_path).encode(), int(threads), int(level))
    if n < 0:
        raise RuntimeError(last_error())
    return n


def align_batch(index, opt_argv, reads1, reads2=None, n_processed=0):
    """Align reads held in host memory. reads*: lists of (name, seq, qual[, comment]).
    Returns (sam_text, stats dict)."""
    L = lib()

    def pack(rs):
        arr = (Read * len(rs))()
        for i, r in enumerate(rs):
            arr[i].name = r[0].encode(); arr[i].seq = r[1].encode()
            arr[i].qual = r[2].encode() if r[2] is not None else None
            arr[i].comment = r[3].encode() if len(r) > 3 and r[3] else None
        return arr
    a1 = pack(reads1)
    a2 = pack(reads2) if reads2 is not None else None
    av = _argv(opt_argv)
    b = L.bsb_batch_create(index._h, len(opt_argv), av, len(reads1), a1, a2)
    if not b:
        raise RuntimeError(last_error())
    try:
        st = RunStats()
        if L.bsb_batch_align(b, int(n_processed), C.byref(st)):
            raise RuntimeError(last_error())
        sam = C.c_char_p(); ln = C.c_size_t()
        if L.bsb_batch_sam(b, C.byref(sam), C.byref(ln), C.byref(st)):
            raise RuntimeError(last_error())
        text = C.string_at(sam, ln.value).decode()
        return text, st.as_dict()
    finally:
        L.bsb_batch_free(b)
