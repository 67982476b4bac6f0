"""TEST INFRASTRUCTURE (oracle) -- ctypes front-end of oracle/liboracle.so (plain-C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (hite_b200/) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INFO = {0: "", 1: "nb", 2: "fl1"}


class HboCall(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("is_te", "info", "copy_num", "align_start", "align_end",
                                       "homo_start", "homo_end", "cons_len", "trim_lead", "trim_trail")]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _mat(rows):
    R = len(rows)
    Cn = len(rows[0]) if R else 0
    assert all(len(r) == Cn for r in rows), "ragged matrix"
    return "".join(rows).encode(), R, Cn


def remove_sparse_cols(rows):
    m, R, Cn = _mat(rows)
    out = C.create_string_buffer(max(1, R * Cn))
    cout = C.c_int(0)
    lib().hbo_remove_sparse_cols(m, R, Cn, out, C.byref(cout))
    k = cout.value
    raw = out.raw
    return [raw[r * k:(r + 1) * k].decode() for r in range(R)]


def judge_v5(cand, rows, plant=1, result_type="cons"):
    """-> dict mirroring judge_boundary_v5's (is_TE, info, cons_seq, copy_num) + boundaries."""
    m, R, Cn = _mat(rows)
    call = HboCall()
    cons = C.create_string_buffer(Cn + 2)
    cb = cand.encode()
    lib().hbo_judge_v5(cb, len(cb), m, R, Cn, int(plant), 1 if result_type == "cons" else 0,
                       C.byref(call), cons)
    return dict(is_te=bool(call.is_te), info=INFO[call.info], cons=cons.value.decode(),
                copy_num=call.copy_num, align_start=call.align_start, align_end=call.align_end,
                homo_start=call.homo_start, homo_end=call.homo_end)


def find_near_matches_ends(p, s, d):
    fs, fe, ls, le = C.c_int(-1), C.c_int(-1), C.c_int(-1), C.c_int(-1)
    n = lib().hbo_find_near_matches(p.encode(), len(p), s.encode(), len(s), d,
                                    C.byref(fs), C.byref(fe), C.byref(ls), C.byref(le))
    return n, (fs.value, fe.value), (ls.value, le.value)


def levenshtein(a, b):
    return lib().hbo_levenshtein(a.encode(), len(a), b.encode(), len(b))
