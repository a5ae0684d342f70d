"""ctypes / numpy mirrors of the PODs declared in include/bkx.h (the C ABI of the hot path).

Plain data layouts only -- no compute.  Shared by the product binding (biokanga_b200.lib) and by the
test-only checker binding under oracle/ so both sides speak the same records.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

NAR_CODES = ["NA", "AA", "EN", "NL", "MH", "ML", "ET", "OJ", "OM", "DP", "DS", "FC", "PR", "UI", "OI", "UP",
             "IS", "IT", "NP", "LC"]  # biokanga/Aligner.cpp:32-51
NAR_COUNT = 20

(NAR_UNALIGNED, NAR_ACCEPTED, NAR_NS, NAR_NOHIT, NAR_MMDELTA, NAR_MULTIALIGN, NAR_TRIM, NAR_SPLICEJCTN,
 NAR_MICROINDEL, NAR_PCRDUP, NAR_NONUNIQUE, NAR_CHROMFILT, NAR_REGIONFILT, NAR_PEINSERTMIN, NAR_PEINSERTMAX,
 NAR_PENOHIT, NAR_PESTRAND, NAR_PECHROM, NAR_PEUNALIGN, NAR_LOCICONSTRAINED) = range(20)

HR_NONE, HR_HITS, HR_MMDELTA, HR_HITINSTS, HR_RMMDELTA = range(5)
PE_NONE, PE_ORPHAN, PE_UNIQUE, PE_ORPHAN_SE, PE_UNIQUE_SE = range(5)
FLG_PE_ALIGNED, FLG_PE_RECOVERED = 1, 2


class Entry(C.Structure):
    _fields_ = [("entry_id", C.c_uint32), ("seq_len", C.c_uint32), ("start_ofs", C.c_uint64),
                ("end_ofs", C.c_uint64), ("name", C.c_char * 88)]


class IndexInfo(C.Structure):
    _fields_ = [("concat_len", C.c_uint64), ("tot_seq_len", C.c_uint64), ("num_entries", C.c_uint32),
                ("sfx_el_size", C.c_uint32), ("version", C.c_uint32), ("attributes", C.c_uint32),
                ("prefix_k", C.c_uint32), ("device", C.c_uint32), ("device_bytes", C.c_uint64),
                ("dataset_name", C.c_char * 84)]


class AlignParams(C.Structure):
    _fields_ = [("pmode", C.c_int32), ("max_subs", C.c_int32), ("min_edit_dist", C.c_int32),
                ("max_ns", C.c_int32), ("align_strand", C.c_int32), ("max_ml_matches", C.c_int32),
                ("min_core_len", C.c_int32), ("max_num_slides", C.c_int32), ("max_iter", C.c_int32),
                ("max_ident_nodes", C.c_int32), ("ml_mode", C.c_int32), ("clamp_max_ml", C.c_int32),
                ("best_matches", C.c_int32), ("reserved", C.c_int32 * 3)]


class PEParams(C.Structure):
    _fields_ = [("pe_proc", C.c_int32), ("pair_min_len", C.c_int32), ("pair_max_len", C.c_int32),
                ("pair_strand", C.c_int32), ("circularised", C.c_int32), ("rescue_core_subs_p1", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class ClusterStats(C.Structure):
    _fields_ = [("multi_reads", C.c_uint32), ("putative", C.c_uint32), ("assigned", C.c_uint32),
                ("near_unique", C.c_uint32), ("near_multi", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class PEStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("unaligned_pairs", "accepted_num_paired", "accepted_num_se",
                                          "partner_paired", "partner_unpaired", "num_filtered_by_chrom",
                                          "under_len_pairs", "over_len_pairs")]


class AlignStats(C.Structure):
    _fields_ = [("nar", C.c_uint64 * NAR_COUNT)] + [(n, C.c_uint64) for n in (
        "plus_hits", "minus_hits", "num_sloughed_ns", "tot_non_aligned", "tot_accepted_unique",
        "tot_accepted_multi", "tot_accepted_aligned", "tot_loci_aligned", "tot_not_accepted_delta",
        "seeds", "cands", "reads")]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n, _ in self._fields_[1:]}
        d["nar"] = [int(v) for v in self.nar]
        return d


RESULT_DTYPE = np.dtype([
    ("nar", "u1"), ("hit_rslt", "u1"), ("strand", "u1"), ("num_hits", "u1"), ("low_mm", "i1"),
    ("nxt_low_mm", "i1"), ("low_hit_instances", "<i2"), ("chrom_id", "<u4"), ("match_loci", "<u4"),
    ("match_len", "<u2"), ("mismatches", "u1"), ("flags", "u1"), ("seeds", "<u4"), ("cands", "<u4"),
    ("reserved", "<u4")])
assert RESULT_DTYPE.itemsize == 32
# bkx_read_result16: the record as it crosses PCIe in the compact host interface
RESULT16_DTYPE = np.dtype([("nar_hr", "u1"), ("strand_flags", "u1"), ("num_hits", "u1"), ("mismatches", "u1"), ("low_mm", "i1"),
                           ("nxt_low_mm", "i1"), ("low_hit_instances", "<i2"), ("chrom_id", "<u4"), ("match_loci", "<u4")])
assert RESULT16_DTYPE.itemsize == 16
MULTI_DTYPE = np.dtype([("chrom_id", "<u4"), ("match_loci", "<u4"), ("match_len", "<u2"), ("strand", "u1"),
                        ("mismatches", "u1")])
assert MULTI_DTYPE.itemsize == 12
assert C.sizeof(Entry) == 112 and C.sizeof(AlignParams) == 64 and C.sizeof(PEParams) == 32

ENTRY_DTYPE = np.dtype([("entry_id", "<u4"), ("seq_len", "<u4"), ("start_ofs", "<u8"), ("end_ofs", "<u8"),
                        ("name", "S88")])
assert ENTRY_DTYPE.itemsize == 112


def ptr(a, ctype=C.c_void_p):
    """ctypes pointer to a numpy array's data."""
    return a.ctypes.data_as(ctype) if ctype is not C.c_void_p else C.c_void_p(a.ctypes.data)
