"""ctypes face of oracle/woltka_oracle.c (CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product
(woltka_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'woltka_oracle.c')
SO = os.path.join(HERE, 'libwoltka_oracle.so')
UNITS = 720720


def build(force=False):
    if not force and os.path.exists(SO) and \
            os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    subprocess.check_call(['gcc', '-O2', '-fopenmp', '-ffp-contract=off',
                           '-shared', '-fPIC', '-o', SO, SRC, '-lm'])
    return SO


class _Plan(C.Structure):
    _fields_ = [('parent', C.c_void_p), ('node_rank', C.c_void_p),
                ('n_nodes', C.c_int32), ('root', C.c_int32),
                ('sub_node', C.c_void_p), ('sub_feat', C.c_void_p),
                ('n_subjects', C.c_int64), ('n_entries', C.c_int32),
                ('kind', C.c_void_p), ('target_rank', C.c_void_p),
                ('flags', C.c_uint32), ('major_th', C.c_double),
                ('subok', C.c_int32), ('n_samples', C.c_int32),
                ('n_features', C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.wko_classify.restype = C.c_int
        _lib.wko_ordinal_match.restype = C.c_int64
        _lib.wko_ordinal_match_naive.restype = C.c_int64
        _lib.wko_match_sweep.restype = C.c_int64
        _lib.wko_max_threads.restype = C.c_int
        _lib.wko_free.argtypes = [C.c_void_p]
    return _lib


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def max_threads():
    return lib().wko_max_threads()


def classify(q, s, *, parent=None, node_rank=None, root=-1, sub_node=None,
             sub_feat=None, kinds, target_rank=None, flags=0, major_th=0.0,
             subok=False, n_samples=1, n_features, q_sample=None,
             q_stratum=None, sample=0, n_threads=1):
    """Returns (units[E,S,NF+1], overflow [(entry, sample, stratum|-1, feature,
    den)], strata {(entry, sample, stratum, feature): units})."""
    q, s = _i32(q), _i32(s)
    keep = [_i32(parent), _i32(node_rank), _i32(sub_node), _i32(sub_feat),
            _i32(kinds), _i32(target_rank if target_rank is not None
                              else np.zeros(len(kinds))),
            _i32(q_sample), _i32(q_stratum)]
    par, nrk, snode, sfeat, knd, trk, qs, qt = keep
    n_subjects = len(sfeat) if sfeat is not None else (
        len(snode) if snode is not None else int(n_features))
    plan = _Plan(_p(par), _p(nrk), 0 if par is None else len(par),
                 -1 if root is None else int(root), _p(snode), _p(sfeat),
                 n_subjects, len(knd), _p(knd), _p(trk), flags,
                 float(major_th), int(bool(subok)), n_samples, n_features)
    units = np.zeros((len(knd), n_samples, n_features + 1), dtype=np.int64)
    ok, od, sk, su = (C.c_void_p() for _ in range(4))
    on, sn = C.c_int64(), C.c_int64()
    rc = lib().wko_classify(
        C.byref(plan), _p(q), _p(s), C.c_int64(len(q)), _p(qs), _p(qt),
        C.c_int32(sample), C.c_int(n_threads), _p(units), C.byref(ok),
        C.byref(od), C.byref(on), C.byref(sk), C.byref(su), C.byref(sn))
    if rc:
        raise ValueError('oracle: subject index out of range')
    ovf_key = np.ctypeslib.as_array(
        C.cast(ok, C.POINTER(C.c_int64)), (max(on.value, 1),))[:on.value].copy()
    ovf_den = np.ctypeslib.as_array(
        C.cast(od, C.POINTER(C.c_int32)), (max(on.value, 1),))[:on.value].copy()
    st_key = np.ctypeslib.as_array(
        C.cast(sk, C.POINTER(C.c_int64)), (max(sn.value, 1),))[:sn.value].copy()
    st_units = np.ctypeslib.as_array(
        C.cast(su, C.POINTER(C.c_int64)), (max(sn.value, 1),))[:sn.value].copy()
    for p_ in (ok, od, sk, su):
        lib().wko_free(p_)
    # canonical, capacity-independent forms
    NF1 = n_features + 1

    def split(key):
        cell = key & ((1 << 40) - 1)
        es, f = divmod(cell, NF1)
        return es // n_samples, es % n_samples, f

    stratified = q_stratum is not None
    strata = {}
    if len(st_key):
        uk, inv = np.unique(st_key, return_inverse=True)
        tot = np.zeros(len(uk), dtype=np.int64)
        np.add.at(tot, inv, st_units)
        for key, val in zip(uk.tolist(), tot.tolist()):
            e, smp, f = split(key)
            strata[(e, smp, key >> 40, f)] = val
    overflow = []
    for key, den in zip(ovf_key.tolist(), ovf_den.tolist()):
        e, smp, f = split(key)
        overflow.append((e, smp, (key >> 40) if stratified else -1, f, den))
    overflow.sort()
    return units, overflow, strata


def _match(fn, contig, beg, end, length, th, contig_off, gbeg, gend):
    cols = [_i32(x) for x in (contig, beg, end, length)]
    contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
    gbeg, gend = _i32(gbeg), _i32(gend)
    cap = max(1024, 4 * len(cols[0]))
    while True:
        r = np.empty(cap, dtype=np.int32)
        g = np.empty(cap, dtype=np.int32)
        n = fn(*[_p(x) for x in cols], C.c_int64(len(cols[0])),
               C.c_double(th), _p(contig_off), _p(gbeg), _p(gend),
               C.c_int32(len(contig_off) - 1), _p(r), _p(g), C.c_int64(cap))
        if n == -1:
            cap *= 4
            continue
        if n < 0:
            raise ValueError('oracle: chunk or contig too large for the '
                             '22-bit index of the reference encoding')
        return r[:n].copy(), g[:n].copy()


def ordinal_match(contig, beg, end, length, th, contig_off, gbeg, gend):
    """(read idx, gene idx) pairs by the reference's sweep, sorted."""
    return _match(lib().wko_ordinal_match, contig, beg, end, length, th,
                  contig_off, gbeg, gend)


def ordinal_match_naive(contig, beg, end, length, th, contig_off, gbeg, gend):
    return _match(lib().wko_ordinal_match_naive, contig, beg, end, length, th,
                  contig_off, gbeg, gend)


def match_sweep(queue, rels):
    """ordinal.match_read_gene on an already merged, sorted code queue."""
    queue = np.ascontiguousarray(queue, dtype=np.int64)
    rels = np.ascontiguousarray(rels, dtype=np.uint32)
    cap = max(1024, 16 * len(queue))
    r = np.empty(cap, dtype=np.int32)
    g = np.empty(cap, dtype=np.int32)
    n = lib().wko_match_sweep(_p(queue), C.c_int64(len(queue)), _p(rels),
                              _p(r), _p(g), C.c_int64(cap))
    if n < 0:
        raise ValueError('oracle: pair buffer too small')
    return list(zip(r[:n].tolist(), g[:n].tolist()))
