"""ctypes binding of libbkx.so -- the C-ABI library that holds the CUDA hot path (include/bkx.h).

This is the only way Python reaches the kernels; it fails loudly when the library or a CUDA device is
missing.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbkx.so")
_LIB = None

EXPORTS = [
    "bkx_abi_version", "bkx_last_error", "bkx_device_count", "bkx_open_index", "bkx_open_index_mem",
    "bkx_open_index_dev", "bkx_clone_index", "bkx_close_index", "bkx_index_info_get", "bkx_get_entry",
    "bkx_get_ident", "bkx_get_seq", "bkx_default_params", "bkx_align_reads", "bkx_align_reads_device",
    "bkx_align_one", "bkx_pair_reads", "bkx_last_kernel_ms", "bkx_kernel_launches",
    "bkx_build_suffix_array_device", "bkx_write_sfx", "bkx_pin_host", "bkx_unpin_host",
    "bkx_pair_reads_device", "bkx_open_index_planes", "bkx_build_suffix_array_planes", "bkx_sort_hits", "bkx_align_reads_packed4",
    "bkx_pack_bases4", "bkx_align_reads_multi", "bkx_align_pairs", "bkx_align_pairs_packed4",
    "bkx_assign_multi_matches", "bkx_self_check", "bkx_debug_reset", "bkx_set_chrom_filter",
    "bkx_align_reads_packed2", "bkx_align_pairs_packed2", "bkx_pack_bases2", "bkx_expand_results16",
    "bkx_align_reads_device_packed2", "bkx_open_index_packed5", "bkx_build_suffix_array_packed5",
    "bkx_alloc_host", "bkx_free_host",
]


class BkxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bkx error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile libbkx.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libbkx.so failed")
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise BkxError(-5, "libbkx.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "-- there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.bkx_abi_version.restype = i32
    L.bkx_last_error.restype = C.c_char_p
    L.bkx_device_count.restype = i32
    L.bkx_open_index.argtypes = [C.c_char_p, i32, i32, C.POINTER(vp)]
    L.bkx_open_index_mem.argtypes = [vp, u64, vp, u32, vp, u32, C.c_char_p, i32, i32, C.POINTER(vp)]
    L.bkx_open_index_dev.argtypes = [vp, u64, vp, u32, vp, u32, C.c_char_p, i32, i32, C.POINTER(vp)]
    L.bkx_open_index_planes.argtypes = [vp, u64, vp, vp, vp, u32, C.c_char_p, i32, i32, C.POINTER(vp)]
    L.bkx_build_suffix_array_planes.argtypes = [vp, u64, vp, vp, i32, u64]
    L.bkx_build_suffix_array_packed5.argtypes = [vp, u64, vp, i32, u64]
    L.bkx_open_index_packed5.argtypes = [vp, u64, vp, vp, u32, C.c_char_p, i32, i32, C.POINTER(vp)]
    L.bkx_sort_hits.argtypes = [vp, u32, vp, i32]
    L.bkx_align_reads_packed4.argtypes = [vp, C.POINTER(abi.AlignParams), vp, vp, u32, vp, C.POINTER(abi.AlignStats)]
    L.bkx_pack_bases4.argtypes = [vp, u64, vp]
    L.bkx_align_reads_packed2.argtypes = [vp, C.POINTER(abi.AlignParams), vp, u64, vp, u32, vp, vp, u64, u32, vp,
                                          C.POINTER(abi.AlignStats)]
    L.bkx_align_pairs_packed2.argtypes = [vp, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), vp, u64, vp, u32, vp, vp, u64, u32,
                                          vp, C.POINTER(abi.AlignStats), C.POINTER(abi.PEStats), vp]
    L.bkx_pack_bases2.argtypes = [vp, u64, vp, vp, vp, u64]
    L.bkx_pack_bases2.restype = C.c_int64
    L.bkx_expand_results16.argtypes = [vp, u32, vp, u32, vp]
    L.bkx_debug_reset.argtypes = [vp, i32]
    L.bkx_self_check.argtypes = [vp]
    L.bkx_self_check.restype = C.c_int64
    L.bkx_assign_multi_matches.argtypes = [vp, u32, vp, i32, i32, u32, C.POINTER(abi.ClusterStats)]
    for fn in (L.bkx_align_pairs, L.bkx_align_pairs_packed4):
        fn.argtypes = [vp, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), vp, vp, u32, vp, C.POINTER(abi.AlignStats),
                       C.POINTER(abi.PEStats), vp]
    L.bkx_align_reads_multi.argtypes = [vp, C.POINTER(abi.AlignParams), vp, vp, u32, vp, vp, C.POINTER(abi.AlignStats)]
    L.bkx_clone_index.argtypes = [vp, i32, C.POINTER(vp)]
    L.bkx_close_index.argtypes = [vp]
    L.bkx_close_index.restype = None
    L.bkx_index_info_get.argtypes = [vp, C.POINTER(abi.IndexInfo)]
    L.bkx_get_entry.argtypes = [vp, u32, C.POINTER(abi.Entry)]
    L.bkx_get_ident.argtypes = [vp, C.c_char_p]
    L.bkx_get_seq.argtypes = [vp, u32, u64, u64, vp]
    L.bkx_get_seq.restype = C.c_int64
    L.bkx_default_params.argtypes = [vp, i32, C.POINTER(abi.AlignParams)]
    L.bkx_align_reads.argtypes = [vp, C.POINTER(abi.AlignParams), vp, vp, u32, vp, C.POINTER(abi.AlignStats)]
    L.bkx_align_reads_device.argtypes = [vp, C.POINTER(abi.AlignParams), vp, vp, u32, u32, vp, vp, vp]
    L.bkx_align_reads_device_packed2.argtypes = [vp, C.POINTER(abi.AlignParams), vp, vp, vp, vp, u32, u32, vp, vp, vp]
    L.bkx_align_one.argtypes = [vp, C.POINTER(abi.AlignParams), vp, i32, C.POINTER(i32), C.POINTER(i32),
                                C.POINTER(i32), vp]
    L.bkx_pair_reads.argtypes = [vp, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), vp, u32, vp, vp,
                                 C.POINTER(abi.PEStats), vp]
    L.bkx_build_suffix_array_device.argtypes = [vp, u64, vp, i32]
    L.bkx_write_sfx.argtypes = [C.c_char_p, vp, u64, vp, u32, vp, u32, C.c_char_p]
    L.bkx_pair_reads_device.argtypes = [vp, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), vp, u32, vp, vp, u32, vp,
                                        vp, vp]
    L.bkx_set_chrom_filter.argtypes = [vp, vp, u32]
    L.bkx_pin_host.argtypes = [vp, C.c_size_t]
    L.bkx_alloc_host.argtypes = [C.c_size_t]
    L.bkx_alloc_host.restype = vp
    L.bkx_free_host.argtypes = [vp]
    L.bkx_free_host.restype = None
    L.bkx_unpin_host.argtypes = [vp]
    L.bkx_last_kernel_ms.argtypes = [vp]
    L.bkx_last_kernel_ms.restype = C.c_float
    L.bkx_kernel_launches.argtypes = [vp]
    L.bkx_kernel_launches.restype = u64
    _LIB = L
    return L


def check(rc):
    if rc < 0:
        raise BkxError(rc, lib().bkx_last_error().decode(errors="replace"))
    return rc


def build_suffix_array_device(d_seq_ptr, concat_len, d_sa_ptr, device=0):
    """GPU suffix-array construction (device pointers; u32 elements out)."""
    check(lib().bkx_build_suffix_array_device(d_seq_ptr, concat_len, d_sa_ptr, device))


def build_suffix_array_planes(d_seq_ptr, concat_len, d_sa_lo_ptr, d_sa_hi_ptr=None, device=0, max_batch=0):
    """Bounded-memory builder for any size (the one for >= 4e9 symbols): u32 low plane + u8 high plane out."""
    check(lib().bkx_build_suffix_array_planes(d_seq_ptr, concat_len, d_sa_lo_ptr, d_sa_hi_ptr, device, max_batch))


def build_suffix_array_packed5(d_seq_ptr, concat_len, d_sa5_ptr, device=0, max_batch=0):
    """Same builder, 5-byte elements back to back out (allocate concat_len * 5 + 16 bytes)."""
    check(lib().bkx_build_suffix_array_packed5(d_seq_ptr, concat_len, d_sa5_ptr, device, max_batch))


def pack_bases4(bases):
    """One-byte base codes -> 4-bit packed (two bases per byte, even base in the low nibble)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    out = np.zeros((bases.size + 1) // 2, dtype=np.uint8)
    check(lib().bkx_pack_bases4(bases.ctypes.data, bases.size, out.ctypes.data))
    return out


def pack_bases2(bases):
    """One-byte base codes -> (2-bit stream [+ 8 bytes of slack], exception positions u64, exception codes u8)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    out = np.zeros((bases.size + 3) // 4 + 8, dtype=np.uint8)
    cap = int((bases & 7 > 3).sum()) + 1
    pos = np.zeros(cap, dtype=np.uint64)
    code = np.zeros(cap, dtype=np.uint8)
    n = check(lib().bkx_pack_bases2(bases.ctypes.data, bases.size, out.ctypes.data, pos.ctypes.data, code.ctypes.data, cap))
    return out, pos[:n].copy(), code[:n].copy()


def expand_results16(res16, lens=None, fixed_len=0):
    """16-byte records -> 32-byte records (seeds / cands 0)."""
    res16 = np.ascontiguousarray(res16, dtype=abi.RESULT16_DTYPE)
    out = np.zeros(len(res16), dtype=abi.RESULT_DTYPE)
    lp = np.ascontiguousarray(lens, dtype=np.uint16) if lens is not None else None
    check(lib().bkx_expand_results16(res16.ctypes.data, len(res16), lp.ctypes.data if lp is not None else None, fixed_len,
                                     out.ctypes.data))
    return out


def assign_multi_matches(results, multi, ml_mode, max_read_len):
    """-r3 / -r4 clustering (host code): rewrites `results` in place, returns the counters of the reference's log."""
    assert results.dtype == abi.RESULT_DTYPE and results.flags.c_contiguous and multi.flags.c_contiguous
    cs = abi.ClusterStats()
    check(lib().bkx_assign_multi_matches(results.ctypes.data, len(results), multi.ctypes.data, multi.shape[1], ml_mode,
                                         max_read_len, C.byref(cs)))
    return cs


def sort_hits(results, device=0):
    """Record indices in the reference's output order (SortHitMatch), ties by index."""
    results = np.ascontiguousarray(results, dtype=abi.RESULT_DTYPE)
    order = np.empty(len(results), dtype=np.uint32)
    check(lib().bkx_sort_hits(results.ctypes.data, len(results), order.ctypes.data, device))
    return order


def write_sfx(path, seq, sa, el_size, entries, name="bkx"):
    """Write a version-5 .sfx file (the container `biokanga index` produces)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    sa = np.ascontiguousarray(sa)
    entries = np.ascontiguousarray(entries, dtype=abi.ENTRY_DTYPE)
    check(lib().bkx_write_sfx(os.fsencode(path), seq.ctypes.data, seq.size, sa.ctypes.data, el_size,
                              entries.ctypes.data, len(entries), name.encode()))


class Index:
    """Device-resident index: the CSfxArrayV3 stand-in (Open/SetTargBlock/getters), SfxArrayV2.h:488-990."""

    def __init__(self, handle):
        self._h = handle
        self.info = abi.IndexInfo()
        check(lib().bkx_index_info_get(self._h, C.byref(self.info)))

    @classmethod
    def open(cls, sfx_path, device=0, prefix_k=0):
        h = C.c_void_p()
        check(lib().bkx_open_index(os.fsencode(sfx_path), device, prefix_k, C.byref(h)))
        return cls(h)

    @classmethod
    def from_host(cls, seq, sa, el_size, entries, name="mem", device=0, prefix_k=0):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        sa = np.ascontiguousarray(sa)
        entries = np.ascontiguousarray(entries, dtype=abi.ENTRY_DTYPE)
        h = C.c_void_p()
        check(lib().bkx_open_index_mem(seq.ctypes.data, seq.size, sa.ctypes.data, el_size, entries.ctypes.data,
                                       len(entries), name.encode(), device, prefix_k, C.byref(h)))
        return cls(h)

    @classmethod
    def from_device(cls, d_seq_ptr, concat_len, d_sa_ptr, el_size, entries, name="dev", device=0, prefix_k=0):
        entries = np.ascontiguousarray(entries, dtype=abi.ENTRY_DTYPE)
        h = C.c_void_p()
        check(lib().bkx_open_index_dev(d_seq_ptr, concat_len, d_sa_ptr, el_size, entries.ctypes.data, len(entries),
                                       name.encode(), device, prefix_k, C.byref(h)))
        return cls(h)

    @classmethod
    def from_planes(cls, d_seq_ptr, concat_len, d_sa_lo_ptr, d_sa_hi_ptr, entries, name="bkx", device=0, prefix_k=0):
        """Index over suffix-array planes already on the device; they are borrowed and must outlive the index."""
        entries = np.ascontiguousarray(entries, dtype=abi.ENTRY_DTYPE)
        h = C.c_void_p()
        check(lib().bkx_open_index_planes(d_seq_ptr, concat_len, d_sa_lo_ptr, d_sa_hi_ptr, entries.ctypes.data,
                                          len(entries), name.encode(), device, prefix_k, C.byref(h)))
        return cls(h)

    @classmethod
    def from_packed5(cls, d_seq_ptr, concat_len, d_sa5_ptr, entries, name="bkx", device=0, prefix_k=0):
        """Index over 5-byte suffix elements already on the device (borrowed; 8-byte aligned, 16 bytes of slack)."""
        entries = np.ascontiguousarray(entries, dtype=abi.ENTRY_DTYPE)
        h = C.c_void_p()
        check(lib().bkx_open_index_packed5(d_seq_ptr, concat_len, d_sa5_ptr, entries.ctypes.data, len(entries), name.encode(),
                                           device, prefix_k, C.byref(h)))
        return cls(h)

    def debug_reset(self, what):
        check(lib().bkx_debug_reset(self._h, what))

    def self_check(self):
        """Suffix-array elements outside the prefix-table bucket of their suffix (0 = the device index is sound)."""
        return check(lib().bkx_self_check(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().bkx_close_index(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- getters (GetIdentName / GetSeqLen / GetIdent / GetSeq) ---
    def entry(self, entry_id):
        e = abi.Entry()
        check(lib().bkx_get_entry(self._h, entry_id, C.byref(e)))
        return e

    def entries(self):
        return [self.entry(i) for i in range(1, self.info.num_entries + 1)]

    def ident(self, name):
        return check(lib().bkx_get_ident(self._h, name.encode()))

    def get_seq(self, entry_id, loci, length):
        buf = np.empty(length, dtype=np.uint8)
        n = check(lib().bkx_get_seq(self._h, entry_id, loci, length, buf.ctypes.data))
        return buf[:n]

    def default_params(self, pmode=0, **kw):
        p = abi.AlignParams()
        check(lib().bkx_default_params(self._h, pmode, C.byref(p)))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    # --- the hot path ---
    def align(self, params, bases, offsets, out=None, stats=None):
        """Host buffers in, host records out (H2D + kernels + D2H inside)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if out is None:
            out = np.zeros(n, dtype=abi.RESULT_DTYPE)
        st = stats if stats is not None else abi.AlignStats()
        check(lib().bkx_align_reads(self._h, C.byref(params), bases.ctypes.data, offsets.ctypes.data, n,
                                    out.ctypes.data, C.byref(st)))
        return out, st

    def align_ptr(self, params, bases_ptr, offsets_ptr, n_reads, out_ptr, stats=None):
        """Same, raw host pointers (pinned buffers owned by the caller)."""
        st = C.byref(stats) if stats is not None else None
        check(lib().bkx_align_reads(self._h, C.byref(params), bases_ptr, offsets_ptr, n_reads, out_ptr, st))

    def align_pairs_ptr(self, params, pe, reads_ptr, offsets_ptr, n_pairs, out_ptr, stats=None, pe_stats=None,
                        len_dist_ptr=None, packed=False):
        """Fused align + pair on raw host pointers; reads one byte per base, or 4-bit packed with packed=True."""
        fn = lib().bkx_align_pairs_packed4 if packed else lib().bkx_align_pairs
        check(fn(self._h, C.byref(params), C.byref(pe), reads_ptr, offsets_ptr, n_pairs, out_ptr,
                 C.byref(stats) if stats is not None else None, C.byref(pe_stats) if pe_stats is not None else None,
                 len_dist_ptr))

    def align_pairs(self, params, pe, bases, offsets, packed=False, len_dist=None):
        """Fused align + pair: (records, align stats, PE stats)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=abi.RESULT_DTYPE)
        st, ps = abi.AlignStats(), abi.PEStats()
        self.align_pairs_ptr(params, pe, bases.ctypes.data, offsets.ctypes.data, n // 2, out.ctypes.data, st, ps,
                             len_dist.ctypes.data if len_dist is not None else None, packed)
        return out, st, ps

    def align_multi(self, params, bases, offsets):
        """-r5 (params.ml_mode = 5): (records, loci[n, max_ml_matches], stats)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=abi.RESULT_DTYPE)
        multi = np.zeros((n, params.max_ml_matches), dtype=abi.MULTI_DTYPE)
        st = abi.AlignStats()
        check(lib().bkx_align_reads_multi(self._h, C.byref(params), bases.ctypes.data, offsets.ctypes.data, n,
                                          out.ctypes.data, multi.ctypes.data, C.byref(st)))
        return out, multi, st

    def align_packed4_ptr(self, params, packed_ptr, offsets_ptr, n_reads, out_ptr, stats=None):
        """Host buffers with the reads 4-bit packed (see pack_bases4); offsets count bases."""
        st = C.byref(stats) if stats is not None else None
        check(lib().bkx_align_reads_packed4(self._h, C.byref(params), packed_ptr, offsets_ptr, n_reads, out_ptr, st))

    def align_packed4(self, params, packed, offsets):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(len(offsets) - 1, dtype=abi.RESULT_DTYPE)
        st = abi.AlignStats()
        self.align_packed4_ptr(params, packed.ctypes.data, offsets.ctypes.data, len(offsets) - 1, out.ctypes.data, st)
        return out, st

    def align_packed2_ptr(self, params, packed2_ptr, lens_ptr, fixed_len, exc_pos_ptr, exc_code_ptr, n_exc, n_reads, out16_ptr,
                          stats=None, pe=None, pe_stats=None, len_dist_ptr=None, first_base=0):
        """The compact host interface on raw host pointers: 2 bits per base in, 16-byte records out (n_reads = reads, also
        for paired ends)."""
        st = C.byref(stats) if stats is not None else None
        if pe is None:
            check(lib().bkx_align_reads_packed2(self._h, C.byref(params), packed2_ptr, first_base, lens_ptr, fixed_len, exc_pos_ptr,
                                                exc_code_ptr, n_exc, n_reads, out16_ptr, st))
        else:
            check(lib().bkx_align_pairs_packed2(self._h, C.byref(params), C.byref(pe), packed2_ptr, first_base, lens_ptr, fixed_len,
                                                exc_pos_ptr, exc_code_ptr, n_exc, n_reads // 2, out16_ptr, st,
                                                C.byref(pe_stats) if pe_stats is not None else None, len_dist_ptr))

    def align_packed2(self, params, bases, offsets, pe=None, len_dist=None, fixed=None, first_base=0):
        """Packs `bases` / `offsets` (one byte per base) into the compact layout, aligns, expands the records again:
        (records, stats[, PE stats]).  fixed=None picks the fixed-length form when every read has the same length;
        first_base > 0 puts that many filler bases in front of the stream (a shard of a longer stream)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        lens = np.diff(offsets).astype(np.uint16)
        o0 = int(offsets[0])
        packed, pos, code = pack_bases2(np.concatenate([np.full(first_base, 4, dtype=np.uint8), bases[o0:int(offsets[-1])]]))
        if first_base:   # the filler (Ns) belongs to some other shard: its exceptions are not this call's
            keep = pos >= first_base
            pos, code = pos[keep].copy(), code[keep].copy()
        if fixed is None:
            fixed = n > 0 and bool((lens == lens[0]).all())
        out16 = np.zeros(n, dtype=abi.RESULT16_DTYPE)
        st, ps = abi.AlignStats(), abi.PEStats()
        self.align_packed2_ptr(params, packed.ctypes.data, None if fixed else lens.ctypes.data, int(lens[0]) if fixed else 0,
                               pos.ctypes.data if len(pos) else None, code.ctypes.data if len(code) else None, len(pos), n,
                               out16.ctypes.data, st, pe, ps, len_dist.ctypes.data if len_dist is not None else None,
                               first_base=first_base)
        res = expand_results16(out16, None if fixed else lens, int(lens[0]) if fixed else 0)
        return (res, st) if pe is None else (res, st, ps)

    def align_device(self, params, d_bases, d_offsets, n_reads, max_read_len, d_out, d_stats=None, stream=None):
        """Device pointers, asynchronous on `stream` (a raw cudaStream_t value or None)."""
        check(lib().bkx_align_reads_device(self._h, C.byref(params), d_bases, d_offsets, n_reads, max_read_len,
                                           d_out, d_stats, stream))

    def align_device_packed2(self, params, d_bases, d_packed2, d_read_flags, d_offsets, n_reads, max_read_len, d_out,
                             d_stats=None, stream=None):
        """Device pointers; the reads are resident one byte per base AND 2-bit packed (see include/bkx.h)."""
        check(lib().bkx_align_reads_device_packed2(self._h, C.byref(params), d_bases, d_packed2, d_read_flags, d_offsets, n_reads,
                                                   max_read_len, d_out, d_stats, stream))

    def align_one(self, params, probe):
        probe = np.ascontiguousarray(probe, dtype=np.uint8)
        hit = np.zeros(1, dtype=abi.RESULT_DTYPE)
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        hr = check(lib().bkx_align_one(self._h, C.byref(params), probe.ctypes.data, len(probe), C.byref(a),
                                       C.byref(b), C.byref(c), hit.ctypes.data))
        return hr, (a.value, b.value, c.value), hit[0]

    def set_chrom_filter(self, keep):
        """-Z / -z inside the pairing (AcceptThisChromID, Aligner.cpp:2651-2710): keep[id] != 0 for the chromosomes whose
        alignments stay, one uint8 per entry id 0..num_entries (index 0 unused); None clears the filter."""
        if keep is None:
            check(lib().bkx_set_chrom_filter(self._h, None, 0))
            return
        kp = np.ascontiguousarray(keep, dtype=np.uint8)
        check(lib().bkx_set_chrom_filter(self._h, kp.ctypes.data, len(kp)))

    def pair(self, params, pe, results, bases=None, offsets=None, len_dist=None):
        n_pairs = len(results) // 2
        st = abi.PEStats()
        b = np.ascontiguousarray(bases, dtype=np.uint8).ctypes.data if bases is not None else None
        o = np.ascontiguousarray(offsets, dtype=np.uint64).ctypes.data if offsets is not None else None
        ld = len_dist.ctypes.data if len_dist is not None else None
        check(lib().bkx_pair_reads(self._h, C.byref(params), C.byref(pe), results.ctypes.data, n_pairs, b, o,
                                   C.byref(st), ld))
        return st

    def pair_device(self, params, pe, d_results, n_pairs, d_bases, d_offsets, max_read_len, d_stats=None,
                    d_len_dist=None, stream=None):
        """Device pointers, asynchronous on `stream`."""
        check(lib().bkx_pair_reads_device(self._h, C.byref(params), C.byref(pe), d_results, n_pairs, d_bases, d_offsets,
                                          max_read_len, d_stats, d_len_dist, stream))

    def last_kernel_ms(self):
        return float(lib().bkx_last_kernel_ms(self._h))

    def kernel_launches(self):
        return int(lib().bkx_kernel_launches(self._h))
