"""CPU check of the two-pairs-per-lane extension kernel (hite_b200/csrc/hb_extend16.cuh).

The warp body is compiled for the host and run by a 32-thread lock-step emulator (tests/emu/),
then compared with the oracle's banded alignment (oracle/hbo_align.c::hbo_align_pair) on seeded
pairs: score, end points and the per-query-base edit records must be identical.  No GPU involved;
the device build of the same source is covered by the `-m gpu` parity tests.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_SRC = os.path.join(HERE, "emu", "emu_extend16.cu")
EMU_SO = os.path.join(HERE, "emu", "_build", "libemu16.so")
HDR = os.path.join(HERE, "..", "hite_b200", "csrc", "hb_extend16.cuh")


def _emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available: cannot build the host emulator")
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    if (not os.path.exists(EMU_SO)) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(EMU_SRC), os.path.getmtime(HDR)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-diag-suppress", "20014", "-Xcompiler", "-fPIC,-pthread", "-shared",
                        "-o", EMU_SO, EMU_SRC], check=True, capture_output=True)
    return C.CDLL(EMU_SO)


class _Aln(C.Structure):
    _fields_ = [("score", C.c_int32), ("i_s", C.c_int32), ("i_e", C.c_int32), ("pad", C.c_int32),
                ("j_s", C.c_int64), ("j_e", C.c_int64)]


CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def _codes(seq):
    return np.array([CODE.get(b, 4) for b in seq.upper().encode()], dtype=np.uint8)


def _revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def _mutate(rng, s, sub, indel, max_run=3):
    out = []
    i = 0
    while i < len(s):
        r = rng.random()
        if r < indel / 2:
            i += int(rng.integers(1, max_run + 1))            # deletion
            continue
        if r < indel:
            out.append("".join(rng.choice(list("ACGT"), int(rng.integers(1, max_run + 1)))))
        c = s[i]
        if rng.random() < sub:
            c = rng.choice([x for x in "ACGT" if x != c])
        out.append(c)
        i += 1
    return "".join(out)


def _case(seed, n_fam, lens, sub, indel, n_runs=0):
    rng = np.random.default_rng(seed)
    genome = list("".join(rng.choice(list("ACGT"), 3000)))
    cands, plan = [], []
    for f in range(n_fam):
        L = int(lens[f % len(lens)])
        fam = "".join(rng.choice(list("ACGT"), L))
        if n_runs:
            p = int(rng.integers(0, max(1, L - 20)))
            fam = fam[:p] + "N" * int(rng.integers(1, 12)) + fam[p:]
        cands.append(fam)
        for _ in range(int(rng.integers(1, 5))):
            copy = _mutate(rng, fam, sub, indel)
            strand = int(rng.integers(0, 2))
            pos = len(genome) + int(rng.integers(20, 200))
            genome.extend(rng.choice(list("ACGT"), pos - len(genome)))
            genome.extend(copy if strand == 0 else _revcomp(copy))
            plan.append((f, strand, pos, len(copy)))
    genome.extend(rng.choice(list("ACGT"), 300))
    return "".join(genome), cands, plan, rng


def _run(seed, n_fam, lens, sub, indel, n_runs=0, edge=False, skew=6, pin=None):
    emu = _emu()
    lib = O.lib()
    genome, cands, plan, rng = _case(seed, n_fam, lens, sub, indel, n_runs)
    ga = genome.encode()
    off = (C.c_uint64 * 2)(0, len(ga))
    lib.hbo_genome_create.restype = C.c_void_p
    g = lib.hbo_genome_create(C.c_char_p(ga), off, C.c_uint32(1))
    lib.hbo_genome_total.restype = C.c_uint64
    tot = lib.hbo_genome_total(C.c_void_p(g))
    gcodes = np.full(tot, 4, dtype=np.uint8)
    gcodes[64:64 + len(ga)] = _codes(genome)
    ccodes = [_codes(c) for c in cands]
    coff = np.zeros(len(cands) + 1, dtype=np.uint64)
    coff[1:] = np.cumsum([len(c) for c in ccodes])
    call = np.concatenate(ccodes)
    p_cand, p_rel, p_j0, p_narrow = [], [], [], []
    for (f, strand, pos, clen) in plan:
        gpos = 64 + pos
        j0 = gpos if strand == 0 else tot - (gpos + clen)
        narrow = int(rng.integers(0, 3))                     # band class: 0 wide (half band 32), 1 narrow (16), 2 tight (8)
        hb = (32, 16, 8)[narrow]
        # seeds put the band roughly, not exactly, on the copy; a large skew runs the path along the band edge
        j0 += int(rng.integers(-skew, skew + 1)) * hb // 16 if skew > 6 else int(rng.integers(-skew, skew + 1)) * min(hb, 16) // 16
        if pin is not None:                                  # put the copy on a chosen diagonal o = pin (scaled to the band)
            j0 = (gpos if strand == 0 else tot - (gpos + clen)) - (pin * hb) // 16
        p_cand.append(f); p_rel.append(strand); p_j0.append(j0); p_narrow.append(narrow)
    if edge:                                                 # windows hanging over both genome ends, and a pure-noise pair
        p_cand += [0, 0, 1, 0, 1]; p_rel += [0, 1, 0, 0, 1]; p_j0 += [-40, tot - 30, 500, -20, tot - 12]; p_narrow += [1, 0, 1, 2, 2]
    n = len(p_cand)
    p_cand = np.array(p_cand, dtype=np.uint32); p_rel = np.array(p_rel, dtype=np.uint32)
    p_j0 = np.array(p_j0, dtype=np.int64); p_narrow = np.array(p_narrow, dtype=np.uint32)
    lens_p = np.array([len(cands[c]) for c in p_cand], dtype=np.uint64)
    qoff = np.zeros(n + 1, dtype=np.uint64); qoff[1:] = np.cumsum(lens_p)
    score = np.zeros(n, dtype=np.int32); i_s = np.zeros(n, dtype=np.int32); i_e = np.zeros(n, dtype=np.int32)
    j_s = np.zeros(n, dtype=np.int64); j_e = np.zeros(n, dtype=np.int64)
    qrec = np.zeros(int(qoff[-1]) + 1, dtype=np.uint32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu_extend16(P(gcodes), C.c_uint64(tot), P(call), P(coff), C.c_uint32(len(cands)), P(p_cand), P(p_rel),
                          P(p_j0), P(p_narrow), C.c_uint32(n), P(score), P(i_s), P(i_e), P(j_s), P(j_e), P(qrec), P(qoff))
    assert rc == 0
    n_pos = 0
    for i in range(n):
        L = int(lens_p[i])
        a = _Aln()
        oq = np.zeros(L, dtype=np.uint32)
        lib.hbo_align_pair(C.c_void_p(g), P(ccodes[p_cand[i]]), C.c_int(L), C.c_int(int(p_rel[i])), C.c_int64(int(p_j0[i])),
                           C.c_int((32, 16, 8)[int(p_narrow[i])]), C.byref(a), P(oq))
        assert score[i] == a.score, (seed, i)
        if a.score > 0:
            n_pos += 1
            assert (i_s[i], i_e[i], j_s[i], j_e[i]) == (a.i_s, a.i_e, a.j_s, a.j_e), (seed, i)
            got = qrec[int(qoff[i]):int(qoff[i]) + L]
            assert np.array_equal(got, oq), (seed, i, np.nonzero(got != oq)[0][:8])
        else:
            assert i_s[i] == -1 and i_e[i] == -1
    lib.hbo_genome_free(C.c_void_p(g))
    return n_pos


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emulated_warp_matches_oracle_mixed_lengths(seed):
    # ragged groups: lengths differ inside a lane-pair and across the lane-groups of one warp
    n_pos = _run(seed, n_fam=5, lens=[90, 333, 150, 64, 401], sub=0.05, indel=0.02)
    assert n_pos >= 5


def test_emulated_warp_indel_heavy_and_n_runs():
    assert _run(11, n_fam=4, lens=[260, 180], sub=0.08, indel=0.06, n_runs=1) >= 4


def test_emulated_warp_genome_edges_and_noise():
    _run(21, n_fam=3, lens=[120, 200, 77], sub=0.02, indel=0.01, edge=True)


@pytest.mark.parametrize("seed", [31, 32, 33, 34])
def test_emulated_warp_path_along_band_edge(seed):
    # copies shifted by up to 15 (narrow) / 30 (wide) diagonals: the best path hugs, or falls off, the band edge
    _run(seed, n_fam=4, lens=[140, 96, 210], sub=0.04, indel=0.04, skew=15)


@pytest.mark.parametrize("pin", [-16, -15, -14, 13, 14, 15])
def test_emulated_warp_copy_pinned_on_edge_diagonals(pin):
    # gaps opened on the outermost band columns exercise the "neighbour outside the band" masks
    _run(40 + pin, n_fam=6, lens=[110, 90, 130], sub=0.02, indel=0.10, pin=pin)
