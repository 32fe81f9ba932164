"""ctypes views of (a) the compiled reference (oracle/_ref/libbwa_ref.so) and (b) the oracle restatement
(oracle/libbsb_oracle.so). Test infrastructure only."""
import ctypes as C
import gzip
import os
import shutil
import struct
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF_SO = os.path.join(ROOT, 'oracle', '_ref', 'libbwa_ref.so')
ORACLE_SO = os.path.join(ROOT, 'oracle', 'libbsb_oracle.so')
_CODE = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'a': 0, 'c': 1, 'g': 2, 't': 3}


def unpack_index(dst=None):
    dst = dst or tempfile.mkdtemp(prefix='bsb_idx_')
    for f in os.listdir(os.path.join(GOLDEN, 'db')):
        with gzip.open(os.path.join(GOLDEN, 'db', f), 'rb') as i, open(os.path.join(dst, f[:-3]), 'wb') as o:
            shutil.copyfileobj(i, o)
    return dst


class Intv(C.Structure):
    _fields_ = [('x', C.c_uint64 * 3), ('info', C.c_uint64)]

    def tolist(self):
        return [int(self.x[0]), int(self.x[1]), int(self.x[2]), int(self.info)]


class IntvV(C.Structure):
    _fields_ = [('n', C.c_size_t), ('m', C.c_size_t), ('a', C.POINTER(Intv))]


def scmat(a, b):
    m = []
    for i in range(4):
        for j in range(4):
            m.append(a if i == j else -b)
        m.append(-1)
    m += [-1] * 5
    return m


class _Common:
    code = staticmethod(lambda c: _CODE.get(c, 4))
    scmat = staticmethod(scmat)


class RefLib(_Common):
    """The reference's own functions."""

    def __init__(self, idxbase):
        L = C.CDLL(REF_SO)
        self.L = L
        L.bwt_restore_bwt.restype = C.c_void_p
        L.bwt_restore_bwt.argtypes = [C.c_char_p]
        L.bwt_restore_sa.argtypes = [C.c_char_p, C.c_void_p]
        self.bwt = L.bwt_restore_bwt((idxbase + '.bwt').encode())
        L.bwt_restore_sa((idxbase + '.sa').encode(), self.bwt)
        hdr = struct.unpack('<5Q', open(idxbase + '.bwt', 'rb').read(40))
        self.primary, self.L2 = hdr[0], [0] + list(hdr[1:])
        self.seq_len = self.L2[4]
        L.bwt_occ4.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.bwt_extend.argtypes = [C.c_void_p, C.POINTER(Intv), C.POINTER(Intv), C.c_int]
        L.bwt_smem1.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(IntvV), C.c_void_p]
        L.bwt_seed_strategy1.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(Intv)]
        L.bwt_sa.restype = C.c_uint64
        L.bwt_sa.argtypes = [C.c_void_p, C.c_uint64]

    def occ4(self, k):
        cnt = (C.c_uint64 * 4)()
        self.L.bwt_occ4(self.bwt, k, cnt)
        return [int(c) for c in cnt]

    def set_intv(self, c):
        return [self.L2[c] + 1, self.L2[3 - c] + 1, self.L2[c + 1] - self.L2[c], 0]

    def extend(self, ik, is_back):
        a = Intv(); a.x[0], a.x[1], a.x[2], a.info = ik
        ok = (Intv * 4)()
        self.L.bwt_extend(self.bwt, C.byref(a), ok, is_back)
        return [[int(o.x[0]), int(o.x[1]), int(o.x[2])] for o in ok]

    def smem1(self, q, x, min_intv):
        v = IntvV()
        ret = self.L.bwt_smem1(self.bwt, len(q), bytes(q), x, min_intv, C.byref(v), None)
        return dict(ret=ret, mem=[v.a[i].tolist() for i in range(v.n)])

    def seed_strategy1(self, q, x, min_len, max_intv):
        m = Intv()
        ret = self.L.bwt_seed_strategy1(self.bwt, len(q), bytes(q), x, min_len, max_intv, C.byref(m))
        return dict(ret=ret, mem=m.tolist())

    def sa(self, k):
        return int(self.L.bwt_sa(self.bwt, k))

    def ksw_extend2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
        r = [C.c_int() for _ in range(5)]
        m = (C.c_int8 * 25)(*mat)
        sc = self.L.ksw_extend2(len(q), bytes(q), len(t), bytes(t), 5, m, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                                *[C.byref(x) for x in r])
        return [sc] + [x.value for x in r]

    def ksw_global2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w):
        m = (C.c_int8 * 25)(*mat)
        n = C.c_int(); cig = C.POINTER(C.c_uint32)()
        sc = self.L.ksw_global2(len(q), bytes(q), len(t), bytes(t), 5, m, o_del, e_del, o_ins, e_ins, w, C.byref(n), C.byref(cig))
        return [sc, [int(cig[i]) for i in range(n.value)]]


class OIndex(C.Structure):
    _fields_ = [('bwt', C.c_void_p), ('sa', C.c_void_p), ('primary', C.c_uint64), ('L2', C.c_uint64 * 5),
                ('seq_len', C.c_uint64), ('sa_intv', C.c_int)]


class OracleLib(_Common):
    """oracle/bsb_oracle.c"""

    def __init__(self, idxbase):
        L = C.CDLL(ORACLE_SO)
        self.L = L
        raw = open(idxbase + '.bwt', 'rb').read()
        hdr = struct.unpack('<5Q', raw[:40])
        self.primary, self.L2 = hdr[0], [0] + list(hdr[1:])
        self.seq_len = self.L2[4]
        self._bwt = np.frombuffer(raw[40:] + b'\0' * 128, dtype=np.uint32).copy()
        sraw = open(idxbase + '.sa', 'rb').read()
        sa_intv, seq_len = struct.unpack('<2Q', sraw[40:56])
        self._sa = np.concatenate([np.array([2 ** 64 - 1], dtype=np.uint64), np.frombuffer(sraw[56:], dtype=np.uint64)])
        ix = OIndex()
        ix.bwt = self._bwt.ctypes.data; ix.sa = self._sa.ctypes.data; ix.primary = self.primary
        for i in range(5):
            ix.L2[i] = self.L2[i]
        ix.seq_len = self.seq_len; ix.sa_intv = int(sa_intv)
        self.ix = ix
        L.bso_occ4.argtypes = [C.POINTER(OIndex), C.c_uint64, C.POINTER(C.c_uint64)]
        L.bso_extend.argtypes = [C.POINTER(OIndex), C.POINTER(Intv), C.POINTER(Intv), C.c_int]
        L.bso_smem.argtypes = [C.POINTER(OIndex), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(Intv), C.POINTER(C.c_int), C.c_int]
        L.bso_seed_forward.argtypes = [C.POINTER(OIndex), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(Intv)]
        L.bso_sa.restype = C.c_uint64
        L.bso_sa.argtypes = [C.POINTER(OIndex), C.c_uint64]

    set_intv = RefLib.set_intv

    def occ4(self, k):
        cnt = (C.c_uint64 * 4)()
        self.L.bso_occ4(C.byref(self.ix), k, cnt)
        return [int(c) for c in cnt]

    def extend(self, ik, is_back):
        a = Intv(); a.x[0], a.x[1], a.x[2], a.info = ik
        ok = (Intv * 4)()
        self.L.bso_extend(C.byref(self.ix), C.byref(a), ok, is_back)
        return [[int(o.x[0]), int(o.x[1]), int(o.x[2])] for o in ok]

    def smem1(self, q, x, min_intv):
        mem = (Intv * (len(q) + 2))(); n = C.c_int()
        ret = self.L.bso_smem(C.byref(self.ix), len(q), bytes(q), x, min_intv, mem, C.byref(n), len(q) + 2)
        return dict(ret=ret, mem=[mem[i].tolist() for i in range(n.value)])

    def seed_strategy1(self, q, x, min_len, max_intv):
        m = Intv()
        ret = self.L.bso_seed_forward(C.byref(self.ix), len(q), bytes(q), x, min_len, max_intv, C.byref(m))
        return dict(ret=ret, mem=m.tolist())

    def sa(self, k):
        return int(self.L.bso_sa(C.byref(self.ix), k))

    def ksw_extend2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
        r = [C.c_int() for _ in range(5)]
        m = (C.c_int8 * 25)(*mat)
        sc = self.L.bso_extend2(len(q), bytes(q), len(t), bytes(t), m, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                                *[C.byref(x) for x in r])
        return [sc] + [x.value for x in r]

    def ksw_global2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w):
        m = (C.c_int8 * 25)(*mat)
        n = C.c_int(); cig = (C.c_uint32 * (len(q) + len(t) + 4))()
        sc = self.L.bso_global2(len(q), bytes(q), len(t), bytes(t), m, o_del, e_del, o_ins, e_ins, w, C.byref(n), cig, len(cig))
        return [sc, [int(cig[i]) for i in range(n.value)]]
