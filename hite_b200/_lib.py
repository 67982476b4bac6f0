"""ctypes binding of libhite_b200.so (include/hite_b200.h).  Fails loudly when the extension is
missing or no GPU is present -- there is no CPU fallback on the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TE_TYPES = {"tir": 0, "helitron": 1, "non_ltr": 2}
INFO = {0: "", 1: "nb", 2: "fl1"}


class HiteB200Error(RuntimeError):
    pass


class HbCall(C.Structure):
    _fields_ = [("is_te", C.c_int32), ("info", C.c_int32), ("copy_num", C.c_int32),
                ("align_start", C.c_int32), ("align_end", C.c_int32),
                ("homo_start", C.c_int32), ("homo_end", C.c_int32), ("cons_len", C.c_int32),
                ("trim_lead", C.c_int32), ("trim_trail", C.c_int32), ("clean_cols", C.c_int32),
                ("flags", C.c_int32), ("cons_off", C.c_int64)]


CALL_DTYPE = np.dtype([("is_te", "<i4"), ("info", "<i4"), ("copy_num", "<i4"), ("align_start", "<i4"),
                       ("align_end", "<i4"), ("homo_start", "<i4"), ("homo_end", "<i4"), ("cons_len", "<i4"),
                       ("trim_lead", "<i4"), ("trim_trail", "<i4"), ("clean_cols", "<i4"), ("flags", "<i4"),
                       ("cons_off", "<i8")])
assert CALL_DTYPE.itemsize == C.sizeof(HbCall)


class HbTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
                ("k_ms", C.c_float * 8), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches", C.c_uint32), ("pad", C.c_uint32)]


def lib_path():
    return os.path.join(_HERE, "libhite_b200.so")


def load():
    """Load the shared library; raises HiteB200Error when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise HiteB200Error(
            "libhite_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hite_b200/csrc`. There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, u8p, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    lib.hb_version.restype = C.c_int
    lib.hb_device_count.restype = C.c_int
    lib.hb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hb_destroy.argtypes = [vp]
    lib.hb_destroy.restype = None
    lib.hb_last_error.argtypes = [vp]
    lib.hb_last_error.restype = C.c_char_p
    lib.hb_get_timing.argtypes = [vp, C.POINTER(HbTiming)]
    lib.hb_vote_matrices.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp,
                                     C.c_uint64]
    _LIB = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One context per process per GPU (include/hite_b200.h).  Not thread-safe."""

    def __init__(self, device=0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.hb_create(int(device), C.byref(h))
        if rc != 0:
            raise HiteB200Error("hb_create(device=%d) failed (%d): %s" %
                                (device, rc, (self._lib.hb_last_error(None) or b"").decode()))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise HiteB200Error("%s failed (%d): %s" % (what, rc, (self._lib.hb_last_error(self._h) or b"").decode()))

    def timing(self):
        t = HbTiming()
        self._lib.hb_get_timing(self._h, C.byref(t))
        return dict(h2d_ms=t.h2d_ms, kernel_ms=t.kernel_ms, d2h_ms=t.d2h_ms, k_ms=list(t.k_ms),
                    h2d_bytes=t.h2d_bytes, d2h_bytes=t.d2h_bytes, launches=t.launches)

    # ------------------------------------------------------------------------------------
    def vote_matrices(self, cands, matrices, te_type="tir", plant=1, result_type="cons"):
        """Judge pre-aligned matrices (kernel iii).  ``matrices`` is a list of lists of equal-length
        row strings (copies only, '-' gaps) exactly as mafft would have written them; returns one dict
        per candidate mirroring ``judge_boundary_v5``'s ``(is_TE, info, cons_seq, copy_num)`` after
        ``remove_sparse_col_in_align_file`` (reference module/Util.py:10419-10428)."""
        n = len(cands)
        assert n == len(matrices)
        if n == 0:
            return []
        R = np.array([len(m) for m in matrices], dtype=np.int32)
        Cc = np.array([len(m[0]) if len(m) else 0 for m in matrices], dtype=np.int32)
        for m, c in zip(matrices, Cc):
            if any(len(r) != c for r in m):
                raise ValueError("ragged alignment matrix")
        mat_off = np.zeros(n + 1, dtype=np.uint64)
        mat_off[1:] = np.cumsum(R.astype(np.uint64) * Cc.astype(np.uint64))
        rows = np.frombuffer("".join("".join(m) for m in matrices).encode("ascii"), dtype=np.uint8)
        cb = [c.encode("ascii") for c in cands]
        cand_off = np.zeros(n + 1, dtype=np.uint64)
        cand_off[1:] = np.cumsum([len(c) for c in cb])
        cand = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        calls = np.zeros(n, dtype=CALL_DTYPE)
        cons_cap = int(Cc.astype(np.int64).sum())
        cons = np.zeros(max(1, cons_cap), dtype=np.uint8)
        rc = self._lib.hb_vote_matrices(self._h, n, _ptr(rows), _ptr(mat_off), _ptr(R), _ptr(Cc), _ptr(cand),
                                        _ptr(cand_off), TE_TYPES[te_type], int(plant),
                                        1 if result_type == "cons" else 0, _ptr(calls), _ptr(cons), cons_cap)
        self._check(rc, "hb_vote_matrices")
        out = []
        cons_b = cons.tobytes()
        for i in range(n):
            c = calls[i]
            seq = cons_b[int(c["cons_off"]):int(c["cons_off"]) + int(c["cons_len"])].decode("ascii")
            out.append(dict(is_te=bool(c["is_te"]), info=INFO[int(c["info"])], cons=seq, copy_num=int(c["copy_num"]),
                            align_start=int(c["align_start"]), align_end=int(c["align_end"]),
                            homo_start=int(c["homo_start"]), homo_end=int(c["homo_end"]),
                            clean_cols=int(c["clean_cols"])))
        return out
