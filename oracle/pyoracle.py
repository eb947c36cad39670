"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy face of oracle/liboracle.so (as_oracle.c).

May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"

ROW_DTYPE = np.dtype([("pos_id", "<i4"), ("A", "<i4"), ("C", "<i4"), ("G", "<i4"), ("T", "<i4"), ("RD", "<i4"),
                      ("Ars", "<i4"), ("Crs", "<i4"), ("Grs", "<i4"), ("Trs", "<i4")])
CALL_DTYPE = np.dtype([("sample", "<i4"), ("row", "<i4"), ("pos_id", "<i4"), ("ref", "i1"), ("alt", "i1"),
                       ("pad", "i1", (2,)), ("k_fw", "<i4"), ("k_bw", "<i4"), ("FW", "<i4"), ("BW", "<i4"),
                       ("p_fw", "<f8"), ("p_bw", "<f8"), ("q_fw", "<f8"), ("q_bw", "<f8"), ("fisher_p", "<f8")],
                      align=True)
ABSENT = np.uint32(0xFFFFFFFF)

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.aso_kf_gammaq.restype = C.c_double
        L.aso_kf_gammaq.argtypes = [C.c_double, C.c_double]
        L.aso_kf_lgamma.restype = C.c_double
        L.aso_kf_lgamma.argtypes = [C.c_double]
        for name in ("aso_poisson_p", "aso_poisson_q"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_int, C.c_int, C.c_float]
        L.aso_q_at_least.restype = C.c_int
        L.aso_q_at_least.argtypes = [C.c_double, C.c_int]
        L.aso_q_from_p.restype = C.c_double
        L.aso_q_from_p.argtypes = [C.c_double]
        L.aso_fisher.restype = C.c_double
        L.aso_fisher.argtypes = [C.c_int] * 4
        L.aso_thr_as_caller_sees.restype = C.c_float
        L.aso_thr_as_caller_sees.argtypes = [C.c_float]
        L.aso_format_thr_cell.restype = C.c_int
        L.aso_format_thr_cell.argtypes = [C.c_float, C.c_float, C.c_int, C.c_char_p]
        L.aso_format_germ_cell.restype = C.c_int
        L.aso_format_germ_cell.argtypes = [C.c_double, C.c_int, C.c_char_p]
        L.aso_homopolymer.restype = C.c_int
        L.aso_homopolymer.argtypes = [C.c_char_p, C.c_char_p, C.c_char]
        L.aso_noise_estimate.restype = None
        L.aso_noise_estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_float, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.aso_call_variants.restype = C.c_int64
        L.aso_call_variants.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int64]
        L.aso_hash_iteration_order.restype = None
        L.aso_hash_iteration_order.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def dense_to_rows(counts: np.ndarray, pos_id: np.ndarray):
    """counts uint32 [S][P][8] (fw ACGT, bw ACGT; absent = all 0xFFFFFFFF) -> (rows, row_off).

    Rows come out per sample in slot order, which is the file order of an ASEQ file written in
    panel order."""
    S, P, _ = counts.shape
    present = counts[:, :, 0] != ABSENT
    row_off = np.zeros(S + 1, dtype=np.int64)
    row_off[1:] = np.cumsum(present.sum(axis=1))
    rows = np.zeros(int(row_off[-1]), dtype=ROW_DTYPE)
    for s in range(S):
        idx = np.nonzero(present[s])[0]
        c = counts[s, idx].astype(np.int64)
        r = rows[row_off[s]:row_off[s + 1]]
        r["pos_id"] = pos_id[idx]
        for b, name in enumerate("ACGT"):
            r[name] = c[:, b] + c[:, 4 + b]
            r[name + "rs"] = c[:, 4 + b]
        r["RD"] = c.sum(axis=1)
    return rows, row_off


def noise_estimate(rows, row_off, U, c_value, cut):
    S = len(row_off) - 1
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    row_off = np.ascontiguousarray(row_off, dtype=np.int64)
    thr = np.empty((U, 4, 2), dtype=np.float32)
    germ_val = np.empty((U, 4), dtype=np.float64)
    germ_present = np.empty((U, 4), dtype=np.uint8)
    count = np.empty((U, 4), dtype=np.int32)
    nrec = np.empty(U, dtype=np.int32)
    lib().aso_noise_estimate(_p(rows), _p(row_off), S, U, np.float32(c_value), int(cut), _p(thr), _p(germ_val),
                             _p(germ_present), _p(count), _p(nrec))
    return {"thr": thr, "germ_val": germ_val, "germ_present": germ_present, "count": count, "nrec": nrec}


def call_variants(rows, row_off, U, ref, thr, cut, cap=None):
    T = len(row_off) - 1
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    row_off = np.ascontiguousarray(row_off, dtype=np.int64)
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    thr = np.ascontiguousarray(thr, dtype=np.float32)
    if cap is None:
        cap = max(1024, 3 * len(rows))
    out = np.zeros(cap, dtype=CALL_DTYPE)
    n = lib().aso_call_variants(_p(rows), _p(row_off), T, U, _p(ref), _p(thr), int(cut), _p(out), cap)
    if n > cap:
        raise RuntimeError(f"oracle call list overflow: {n} > {cap}")
    return out[:n]


def hash_iteration_order(keys) -> list:
    """Iteration order of a libstdc++ unordered_map<string,string> after inserting `keys` in order
    (EE:1081, VC:672, VC:1046).  Returns indices into `keys` (duplicates collapse to the first)."""
    arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
    out = np.empty(len(keys), dtype=np.int32)
    lib().aso_hash_iteration_order(arr, len(keys), _p(out))
    return [int(i) for i in out if i >= 0]


def kf_gammaq(s, z):
    return lib().aso_kf_gammaq(float(s), float(z))


def poisson_p(k, rd, err):
    return lib().aso_poisson_p(int(k), int(rd), float(np.float32(err)))


def poisson_q(k, rd, err):
    return lib().aso_poisson_q(int(k), int(rd), float(np.float32(err)))


def q_at_least(p, threshold=5) -> bool:
    return bool(lib().aso_q_at_least(float(p), int(threshold)))


def q_from_p(p) -> float:
    return lib().aso_q_from_p(float(p))


def fisher(a, b, c, d):
    return lib().aso_fisher(int(a), int(b), int(c), int(d))


def thr_as_caller_sees(thr: np.ndarray) -> np.ndarray:
    """Element-wise "%f" -> stof round trip of float32 thresholds (EE:1787 -> VC:889-890)."""
    flat = np.ascontiguousarray(thr, dtype=np.float32).ravel()
    out = np.array([lib().aso_thr_as_caller_sees(float(v)) for v in flat], dtype=np.float32)
    return out.reshape(np.shape(thr))


def format_thr_cell(fw, bw, is_ref) -> str:
    buf = C.create_string_buffer(96)
    lib().aso_format_thr_cell(float(fw), float(bw), int(is_ref), buf)
    return buf.value.decode()


def format_germ_cell(v, present) -> str:
    buf = C.create_string_buffer(64)
    lib().aso_format_germ_cell(float(v), int(present), buf)
    return buf.value.decode()


def homopolymer(down: str, up: str, alt: str) -> int:
    return lib().aso_homopolymer(down.encode(), up.encode(), alt.encode())
