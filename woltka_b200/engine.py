"""Thin Python face of one GPU context of the C-ABI (include/woltka_b200.h).

The engine speaks integer SoA only; strings are interned one level up
(woltka_b200.session).  Host inputs are numpy int32 arrays, device inputs are
raw device addresses (e.g. ``tensor.data_ptr()``).
"""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _lib
from ._lib import (UNITS, MAX_ENTRIES, KIND_NONE, KIND_FREE, KIND_RANK,
                   KIND_NONE_ID, F_UNIQ, F_ABOVE, F_MAJOR, F_UNASSIGNED,
                   WoltkaB200Error)

ASSIGN_UNIQ = 1 << 30

__all__ = ['Engine', 'ASSIGN_UNIQ', 'UNITS', 'MAX_ENTRIES', 'KIND_NONE', 'KIND_FREE',
           'KIND_RANK', 'KIND_NONE_ID', 'F_UNIQ', 'F_ABOVE', 'F_MAJOR',
           'F_UNASSIGNED', 'WoltkaB200Error', 'pinned_empty']


def _i32(a):
    """Contiguous int32 view/copy of an array-like, or None."""
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class _Pinned:
    def __init__(self, nbytes):
        self.lib = _lib.load()
        p = C.c_void_p()
        _lib.check(self.lib.wk_host_alloc(C.byref(p), nbytes))
        self.ptr = p.value
        self.nbytes = nbytes

    def __del__(self):
        try:
            self.lib.wk_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(n, dtype=np.int32):
    """numpy array backed by page-locked host memory (wk_host_alloc)."""
    dtype = np.dtype(dtype)
    owner = _Pinned(max(int(n), 1) * dtype.itemsize)
    buf = (C.c_char * owner.nbytes).from_address(owner.ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    arr = arr.view(_PinnedArray)
    arr._owner = owner
    return arr


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, '_owner', None)


def release_cached_memory(device=0):
    """Return the device blocks kept from closed engines to the driver."""
    _lib.check(_lib.load().wk_release_cached_memory(int(device)))


class PackedChunk:
    """A chunk in the compact wire format: head bits (uint64 words) and
    subjects — a uint16 / uint32 array (`width` 16 / 32) or a little-endian bit
    stream of `width` bits per subject held in uint64 words."""

    def __init__(self, bits, subj, n, width=None):
        self.bits, self.subj, self.n = bits, subj, n
        self.width = width or subj.dtype.itemsize * 8
        self.stream = subj.dtype == np.uint64

    @property
    def nbytes(self):
        return (self.n + 63) // 64 * 8 + (self.n * self.width + 7) // 8


# widths the host packer writes: k subjects fill a whole number of bytes of
# one 64-bit word (k = 1, 4, 4, 4, -, 2, 2, 2, -)
_STREAM_WIDTHS = {8: 1, 10: 4, 12: 4, 14: 4, 20: 2, 24: 2, 28: 2}


def _bit_stream(sidx, width, alloc):
    """Subjects as a little-endian bit stream of `width` bits each."""
    n, k = len(sidx), _STREAM_WIDTHS[width]
    groups = (n + k - 1) // k
    v = np.zeros(groups * k, dtype=np.uint64)
    v[:n] = sidx
    v = v.reshape(groups, k)
    word = v[:, 0].copy()
    for j in range(1, k):
        word |= v[:, j] << np.uint64(j * width)
    nb = k * width // 8                      # bytes per group
    out = alloc((groups * nb + 7) // 8 + 2, np.uint64)
    out[:] = 0
    out.view(np.uint8)[:groups * nb] = \
        word.view(np.uint8).reshape(groups, 8)[:, :nb].reshape(-1)
    return out


class Engine:
    """One wk_ctx (one GPU)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        ctx = C.c_void_p()
        _lib.check(self.lib.wk_create(device, C.byref(ctx)))
        self.ctx = ctx
        self.device = device
        self.E = self.S = 0
        self.NF = 0
        self.T = 0

    def close(self):
        if getattr(self, 'ctx', None):
            self.lib.wk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ----------------------------------------------------------
    def set_stream(self, cuda_stream):
        """Run on an existing cudaStream_t handle; 0 means the legacy default
        stream (torch's default), None restores the context's own stream."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1   # cudaStreamLegacy == 0x1
        _lib.check(self.lib.wk_set_stream(self.ctx, C.c_void_p(handle)))

    def sync(self):
        _lib.check(self.lib.wk_sync(self.ctx))

    def launch_count(self):
        return int(self.lib.wk_launch_count(self.ctx))

    def last_kernel(self):
        return self.lib.wk_last_kernel(self.ctx).decode()

    def set_tuning(self, grid=0, block=0, cache_slots=0):
        _lib.check(self.lib.wk_set_tuning(self.ctx, grid, block, cache_slots))

    def set_option(self, name, value):
        """Named knob of the context (wk_set_option); 0 = default."""
        _lib.check(self.lib.wk_set_option(self.ctx, name.encode(), int(value)))

    # -- model -------------------------------------------------------------
    def set_tree(self, parent, root):
        parent = _i32(parent)
        self._keep_parent = parent
        _lib.check(self.lib.wk_set_tree(self.ctx, _ptr(parent), len(parent),
                                        -1 if root is None else int(root)))
        self.T = len(parent)

    def set_plan(self, kinds, flags=0, major_th=0.0, n_samples=1,
                 n_features=0):
        kinds = _i32(kinds)
        _lib.check(self.lib.wk_set_plan(self.ctx, _ptr(kinds), len(kinds),
                                        flags, float(major_th), n_samples,
                                        n_features))
        self.E, self.S, self.NF = len(kinds), n_samples, n_features
        self.kinds = kinds

    def resize_counts(self, n_samples, n_features):
        _lib.check(self.lib.wk_resize_counts(self.ctx, n_samples, n_features))
        self.S, self.NF = n_samples, n_features

    def set_subjects(self, tab, sub_node=None, n_subjects=None):
        tab = _i32(tab)
        sub_node = _i32(sub_node)
        if n_subjects is None:
            n_subjects = (tab.shape[-1] if tab is not None and tab.ndim == 2
                          else len(sub_node) if sub_node is not None else 0)
        if tab is not None and tab.size != self.E * n_subjects:
            raise ValueError('tab must have shape [n_entries, n_subjects]')
        _lib.check(self.lib.wk_set_subjects(self.ctx, _ptr(tab),
                                            _ptr(sub_node), n_subjects))
        self.V = n_subjects

    # -- classify ----------------------------------------------------------
    def classify_chunk(self, qidx, sidx, q_sample=None, q_stratum=None,
                       sample=0):
        qidx, sidx = _i32(qidx), _i32(sidx)
        q_sample, q_stratum = _i32(q_sample), _i32(q_stratum)
        if len(qidx) != len(sidx):
            raise ValueError('qidx and sidx differ in length')
        n_qry = max(len(q_sample) if q_sample is not None else 0,
                    len(q_stratum) if q_stratum is not None else 0)
        _lib.check(self.lib.wk_classify_chunk(
            self.ctx, _ptr(qidx), _ptr(sidx), len(qidx), _ptr(q_sample),
            _ptr(q_stratum), n_qry, sample))

    @staticmethod
    def pack_columns(qidx, sidx, pinned=True, n_subjects=None):
        """The compact wire format of a chunk (wk_classify_packed[_bits]): one
        head bit per record and the subjects in as few bits as the largest
        index (or `n_subjects` - 1) needs: a bit stream of 8 / 10 / 12 / 14 /
        20 / 24 / 28 bits each, else a uint16 / uint32 array; in page-locked
        memory."""
        qidx, sidx = np.asarray(qidx), np.asarray(sidx)
        n = len(qidx)
        heads = np.empty(n, dtype=bool)
        if n:
            heads[0] = True
            np.not_equal(qidx[1:], qidx[:-1], out=heads[1:])
        packed = np.packbits(heads, bitorder='little')
        words = (n + 63) // 64
        alloc = pinned_empty if pinned else (lambda k, dt: np.empty(k, dtype=dt))
        bits = alloc(max(words, 1), np.uint64)
        bits[:] = 0
        bits.view(np.uint8)[:len(packed)] = packed
        top = max(int(sidx.max()) if n else 0, (n_subjects or 1) - 1)
        need = max(top.bit_length(), 1)
        width = min(w for w in (8, 10, 12, 14, 16, 20, 24, 28, 32) if w >= need)
        if width in _STREAM_WIDTHS and n:
            return PackedChunk(bits, _bit_stream(sidx, width, alloc), n, width)
        dt = np.uint16 if width <= 16 else np.uint32
        subj = alloc(max(n, 1), dt)
        subj[:n] = sidx
        return PackedChunk(bits, subj, n)

    def classify_packed(self, packed, q_sample=None, sample=0, q_stratum=None):
        q_sample, q_stratum = _i32(q_sample), _i32(q_stratum)
        n_qry = max(len(q_sample) if q_sample is not None else 0,
                    len(q_stratum) if q_stratum is not None else 0)
        if packed.stream:
            _lib.check(self.lib.wk_classify_packed_bits(
                self.ctx, _ptr(packed.bits), _ptr(packed.subj), packed.width,
                packed.n, _ptr(q_sample), _ptr(q_stratum), n_qry, sample))
            return
        _lib.check(self.lib.wk_classify_packed(
            self.ctx, _ptr(packed.bits), _ptr(packed.subj),
            packed.subj.dtype.itemsize, packed.n, _ptr(q_sample),
            _ptr(q_stratum), n_qry, sample))

    def classify_device(self, d_qidx, d_sidx, n_rec, d_q_sample=None,
                        d_q_stratum=None, n_qry=0, sample=0):
        _lib.check(self.lib.wk_classify_device(
            self.ctx, _ptr(d_qidx), _ptr(d_sidx), n_rec, _ptr(d_q_sample),
            _ptr(d_q_stratum), n_qry, sample))

    # -- ordinal -----------------------------------------------------------
    def ordinal_set_genes(self, contig_off, gbeg, gend, gene_subject):
        contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
        gbeg, gend, gene_subject = _i32(gbeg), _i32(gend), _i32(gene_subject)
        _lib.check(self.lib.wk_ordinal_set_genes(
            self.ctx, _ptr(contig_off), _ptr(gbeg), _ptr(gend),
            _ptr(gene_subject), len(contig_off) - 1, len(gbeg)))

    def ordinal_chunk(self, qidx, contig, beg, end, length, th, q_sample=None,
                      q_stratum=None, sample=0):
        cols = [_i32(x) for x in (qidx, contig, beg, end, length)]
        if len({len(x) for x in cols}) != 1:
            raise ValueError('record columns differ in length')
        q_sample, q_stratum = _i32(q_sample), _i32(q_stratum)
        n_qry = max(len(q_sample) if q_sample is not None else 0,
                    len(q_stratum) if q_stratum is not None else 0)
        _lib.check(self.lib.wk_ordinal_chunk(
            self.ctx, *[_ptr(x) for x in cols], len(cols[0]), float(th),
            _ptr(q_sample), _ptr(q_stratum), n_qry, sample))

    def ordinal_device(self, d_cols, n_rec, th, d_q_sample=None,
                       d_q_stratum=None, n_qry=0, sample=0):
        _lib.check(self.lib.wk_ordinal_device(
            self.ctx, *[_ptr(x) for x in d_cols], n_rec, float(th),
            _ptr(d_q_sample), _ptr(d_q_stratum), n_qry, sample))

    def ordinal_enable_pairs(self):
        n = C.c_int64()
        _lib.check(self.lib.wk_ordinal_fetch_pairs(self.ctx, C.byref(n), None,
                                                   None, 0))
        return n.value

    def ordinal_pairs(self):
        """(read index, gene index) pairs of the last ordinal chunk."""
        n = self.ordinal_enable_pairs()
        r = np.empty(n, dtype=np.int32)
        g = np.empty(n, dtype=np.int32)
        if n:
            m = C.c_int64()
            _lib.check(self.lib.wk_ordinal_fetch_pairs(
                self.ctx, C.byref(m), _ptr(r), _ptr(g), n))
        return r, g

    # -- results -----------------------------------------------------------
    def fetch_counts(self):
        """int64 units table [n_entries, n_samples, n_features + 1]."""
        out = np.empty((self.E, self.S, self.NF + 1), dtype=np.int64)
        _lib.check(self.lib.wk_fetch_counts(self.ctx, _ptr(out)))
        return out

    def fetch_overflow(self):
        """(cell, stratum, den): shares 1/den that do not divide UNITS; cell is
        the flat index into the units table, stratum -1 unless stratified."""
        n = C.c_int64()
        _lib.check(self.lib.wk_fetch_overflow(self.ctx, C.byref(n), None, None,
                                              None, 0))
        cell = np.empty(n.value, dtype=np.int64)
        strat = np.empty(n.value, dtype=np.int32)
        den = np.empty(n.value, dtype=np.int32)
        if n.value:
            _lib.check(self.lib.wk_fetch_overflow(
                self.ctx, C.byref(n), _ptr(cell), _ptr(strat), _ptr(den),
                len(cell)))
        return cell, strat, den

    def fetch_strata(self):
        """(entry, sample, stratum, feature, units) arrays.  Large results
        land in page-locked buffers the engine keeps (D2H at full PCIe rate);
        the arrays are views of them, valid until the next fetch_strata()."""
        n = C.c_int64()
        _lib.check(self.lib.wk_fetch_strata(self.ctx, C.byref(n), None, None,
                                            None, None, None, 0))
        m = n.value
        spec = (('e', np.int32), ('s', np.int32), ('t', np.int32),
                ('f', np.int64), ('u', np.int64))
        if m >= (1 << 20):
            pin = getattr(self, '_pin', None)
            if pin is None or len(pin['e']) < m:
                cap = m + m // 8
                pin = self._pin = {k: pinned_empty(cap, dt) for k, dt in spec}
            out = [pin[k][:m] for k, _ in spec]
        else:
            out = [np.empty(m, dtype=dt) for _, dt in spec]
        if m:
            _lib.check(self.lib.wk_fetch_strata(
                self.ctx, C.byref(n), *[_ptr(x) for x in out], m))
        return tuple(out)

    def set_assign_output(self, enable=True):
        _lib.check(self.lib.wk_set_assign_output(self.ctx, int(enable)))

    def fetch_assignments(self, n_rec):
        """int32 [n_entries, n_rec] of the last plain chunk (read maps): -1,
        feature | ASSIGN_UNIQ on a query's first record, or a list member."""
        out = np.empty((self.E, n_rec), dtype=np.int32)
        _lib.check(self.lib.wk_fetch_assignments(self.ctx, _ptr(out), n_rec))
        return out

    def reset_counts(self):
        _lib.check(self.lib.wk_reset_counts(self.ctx))

    def counts_device(self):
        """(device address, number of int64 elements) of the units table."""
        p = C.c_void_p()
        n = C.c_int64()
        _lib.check(self.lib.wk_counts_device(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- SAM text on the device (wk_parse.cuh) -----------------------------
    FORMATS = {'sam': 0, 'b6o': 1, 'paf': 2, 'map': 3}

    def parse_sam(self, text, demux=False, fmt='sam'):
        """Parse a chunk of alignment text (bytes; SAM body or b6o / paf /
        map lines); returns (n_rec, n_qry, n_subjects_total,
        n_samples_total)."""
        if isinstance(text, np.ndarray):      # e.g. pinned_empty(n, np.uint8)
            buf = np.ascontiguousarray(text, dtype=np.uint8)
            ptr, nbytes = C.c_void_p(buf.ctypes.data), buf.nbytes
        else:
            buf = text if isinstance(text, bytes) else bytes(text)
            ptr, nbytes = C.cast(C.c_char_p(buf), C.c_void_p), len(buf)
        n_rec, n_qry = C.c_int64(), C.c_int64()
        n_sub, n_smp = C.c_int32(), C.c_int32()
        _lib.check(self.lib.wk_parse_text(
            self.ctx, ptr, nbytes,
            self.FORMATS[fmt], int(bool(demux)),
            C.byref(n_rec), C.byref(n_qry), C.byref(n_sub), C.byref(n_smp)))
        return n_rec.value, n_qry.value, n_sub.value, n_smp.value

    def parse_block(self, buf, demux=False, fmt='sam', final=True):
        """Parse one block of a file (uint8 array, best page-locked); unless
        `final`, the last query group and any cut line are left for the next
        block.  Returns (consumed bytes, n_rec, n_qry, n_subjects_total,
        n_samples_total) (wk_parse_block)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        used, n_rec, n_qry = C.c_int64(), C.c_int64(), C.c_int64()
        n_sub, n_smp = C.c_int32(), C.c_int32()
        _lib.check(self.lib.wk_parse_block(
            self.ctx, C.c_void_p(buf.ctypes.data), buf.nbytes,
            self.FORMATS[fmt], int(bool(demux)), int(bool(final)),
            C.byref(used), C.byref(n_rec), C.byref(n_qry), C.byref(n_sub),
            C.byref(n_smp)))
        return used.value, n_rec.value, n_qry.value, n_sub.value, n_smp.value

    def fetch_names(self, which, lo, hi):
        """Interned names [lo, hi) of the subject (0) / sample (1) table."""
        if hi <= lo:
            return []
        cap = 1 << 20
        while True:
            buf = C.create_string_buffer(cap)
            lens = np.empty(hi - lo, dtype=np.int32)
            used = C.c_int64()
            rc = self.lib.wk_parse_fetch_names(self.ctx, which, lo, hi, buf, cap,
                                               C.byref(used), _ptr(lens))
            if rc == 5 and cap < (1 << 30):     # WK_ERR_CAPACITY: grow
                cap *= 8
                continue
            _lib.check(rc)
            break
        raw = buf.raw[:used.value]
        out, at = [], 0
        for n in lens.tolist():
            out.append(raw[at:at + n].decode())
            at += n
        return out

    def fetch_parsed_columns(self, n_rec, n_qry, demux=False):
        q = np.empty(n_rec, dtype=np.int32)
        s = np.empty(n_rec, dtype=np.int32)
        qs = np.empty(n_qry, dtype=np.int32) if demux else None
        ql = np.empty(n_qry, dtype=np.uint32)
        _lib.check(self.lib.wk_parse_fetch_columns(self.ctx, _ptr(q), _ptr(s),
                                                   _ptr(qs), _ptr(ql)))
        return q, s, qs, ql

    def classify_parsed(self, sample_map=None, sample=0):
        sm = _i32(sample_map)
        _lib.check(self.lib.wk_classify_parsed(
            self.ctx, _ptr(sm), 0 if sm is None else len(sm), sample))

    def parse_options(self, trimsub=None, exclude=None, coords=False):
        """Reader options for the chunks that follow (wk_parse_options):
        `--trim-sub` separator, `--exclude` names, coordinates for the
        coordinate matcher."""
        trim = (trimsub or '').encode()
        names = [x.encode() for x in (exclude or ())]
        lens = np.asarray([len(x) for x in names], dtype=np.int32)
        blob = b''.join(names)
        _lib.check(self.lib.wk_parse_options(
            self.ctx, C.cast(C.c_char_p(trim), C.c_void_p), len(trim),
            C.cast(C.c_char_p(blob), C.c_void_p), _ptr(lens), len(names),
            int(bool(coords))))

    def fetch_parsed_coords(self, n_rec):
        cols = [np.empty(n_rec, dtype=np.int32) for _ in range(3)]
        _lib.check(self.lib.wk_parse_fetch_coords(self.ctx,
                                                  *[_ptr(x) for x in cols]))
        return cols

    def ordinal_parsed(self, contig_map, th, sample_map=None, sample=0):
        """Match + classify the last chunk parsed with coordinates."""
        cm = _i32(contig_map)
        sm = _i32(sample_map)
        _lib.check(self.lib.wk_ordinal_parsed(
            self.ctx, _ptr(cm), len(cm), float(th), _ptr(sm),
            0 if sm is None else len(sm), sample))

    # -- subject coverage (--outcov) ------------------------------------------
    def cover_add(self, sample, subject, beg, end):
        """Append intervals [beg, end] of (sample, subject) index pairs
        (range.parse_ranges, range.py:112-151)."""
        cols = [_i32(x) for x in (sample, subject, beg, end)]
        if len({len(x) for x in cols}) != 1:
            raise ValueError('interval columns differ in length')
        _lib.check(self.lib.wk_cover_add(self.ctx, *[_ptr(x) for x in cols],
                                         len(cols[0])))

    def cover_ranges(self):
        """Merged ranges (range.calc_coverage, range.py:154-180) as four
        columns ordered by (sample, subject, beg)."""
        n = C.c_int64()
        _lib.check(self.lib.wk_cover_merge(self.ctx, C.byref(n)))
        out = [np.empty(n.value, dtype=np.int32) for _ in range(4)]
        if n.value:
            _lib.check(self.lib.wk_cover_fetch(self.ctx, *[_ptr(x) for x in out],
                                               n.value))
        return out

    def device_tensor(self, ptr, n, typestr='<i8'):
        """Zero-copy torch view of n elements of this context's device memory."""
        import torch

        class _Wrap:
            __cuda_array_interface__ = {
                'shape': (int(n),), 'typestr': typestr,
                'data': (int(ptr), False), 'version': 2}
        if not n:
            dt = {'<i8': torch.int64, '<i4': torch.int32}[typestr]
            return torch.empty(0, dtype=dt, device=f'cuda:{self.device}')
        return torch.as_tensor(_Wrap(), device=f'cuda:{self.device}')

    def counts_tensor(self):
        """Zero-copy torch view of the units table (for an NCCL reduce)."""
        ptr, n = self.counts_device()
        return self.device_tensor(ptr, n)

    # -- merging contexts (distributed.merge_engine) -----------------------
    def strata_export(self):
        """(keys, units) int64 views of the compacted strata cells; valid
        until the next call on this engine."""
        k, v, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self.lib.wk_strata_export_device(
            self.ctx, C.byref(k), C.byref(v), C.byref(n)))
        return (self.device_tensor(k.value, n.value),
                self.device_tensor(v.value, n.value))

    def strata_reserve(self, n_cells):
        """Room for n_cells more strata cells, in one step."""
        _lib.check(self.lib.wk_strata_reserve(self.ctx, int(n_cells)))

    def reset_strata(self):
        _lib.check(self.lib.wk_reset_strata(self.ctx))

    def strata_import(self, keys, units):
        """Add (key, units) cells (int64 CUDA tensors) into the strata table."""
        _lib.check(self.lib.wk_strata_import_device(
            self.ctx, C.c_void_p(keys.data_ptr()),
            C.c_void_p(units.data_ptr()), keys.numel()))

    def overflow_export(self):
        """(keys int64, den int32) views of the overflow list."""
        k, d, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self.lib.wk_overflow_export_device(
            self.ctx, C.byref(k), C.byref(d), C.byref(n)))
        return (self.device_tensor(k.value, n.value),
                self.device_tensor(d.value, n.value, '<i4'))

    def overflow_import(self, keys, den, stratified=False):
        _lib.check(self.lib.wk_overflow_import_device(
            self.ctx, C.c_void_p(keys.data_ptr()), C.c_void_p(den.data_ptr()),
            keys.numel(), int(bool(stratified))))


def units_to_value(units, extra=None):
    """Exact count from units (+ optional list of overflow denominators):
    int when integral, else the correctly rounded double."""
    if not extra:
        q, r = divmod(int(units), UNITS)
        return q if r == 0 else float(Fraction(int(units), UNITS))
    v = Fraction(int(units), UNITS)
    for d in extra:
        v += Fraction(1, int(d))
    return int(v) if v.denominator == 1 else float(v)
