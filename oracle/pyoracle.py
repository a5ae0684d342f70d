"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_build/libbkoracle.so (the CPU restatement).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from biokanga_b200 import abi  # noqa: E402  (POD layouts only)

_LIB = None
REF_BIN = os.path.join(_HERE, "_ref", "biokanga")


def build():
    """Compile the C restatement (gcc) and, if /root/reference exists, the reference binary."""
    subprocess.run(["make", "-s", "-C", _HERE, "_build/libbkoracle.so"], check=True)
    if os.path.isdir("/root/reference/biokanga"):
        subprocess.run([os.path.join(_HERE, "build_ref.sh")], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libbkoracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.bko_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.bko_open_mem.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                   C.POINTER(C.c_void_p)]
        L.bko_close.argtypes = [C.c_void_p]
        L.bko_info.argtypes = [C.c_void_p, C.POINTER(abi.IndexInfo)]
        L.bko_get_entry.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(abi.Entry)]
        L.bko_default_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.AlignParams)]
        L.bko_locate_first_exact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.bko_locate_first_exact.restype = C.c_int64
        L.bko_locate_last_exact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.bko_locate_last_exact.restype = C.c_int64
        L.bko_align_batch.argtypes = [C.c_void_p, C.POINTER(abi.AlignParams), C.c_void_p, C.c_void_p, C.c_uint32,
                                      C.c_void_p, C.POINTER(abi.AlignStats), C.c_int]
        L.bko_align_batch_multi.argtypes = [C.c_void_p, C.POINTER(abi.AlignParams), C.c_void_p, C.c_void_p, C.c_uint32,
                                            C.c_void_p, C.c_void_p, C.POINTER(abi.AlignStats), C.c_int]
        L.bko_pair_reads.argtypes = [C.c_void_p, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), C.c_void_p,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(abi.PEStats), C.c_void_p]
        L.bko_pair_reads_filtered.argtypes = [C.c_void_p, C.POINTER(abi.AlignParams), C.POINTER(abi.PEParams), C.c_void_p,
                                              C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(abi.PEStats), C.c_void_p, C.c_void_p]
        L.bko_seq.argtypes = [C.c_void_p]
        L.bko_seq.restype = C.c_void_p
        L.bko_sa.argtypes = [C.c_void_p]
        L.bko_sa.restype = C.c_void_p
        _LIB = L
    return _LIB


class OracleIndex:
    def __init__(self, sfx_path=None, *, seq=None, sa=None, el_size=4, entries=None):
        self._h = C.c_void_p()
        L = lib()
        if sfx_path is not None:
            rc = L.bko_open(os.fsencode(sfx_path), C.byref(self._h))
        else:
            self._keep = (seq, sa, entries)
            rc = L.bko_open_mem(seq.ctypes.data, seq.size, sa.ctypes.data, el_size, entries.ctypes.data,
                                len(entries), C.byref(self._h))
        if rc < 0:
            raise RuntimeError("bko_open failed: %d" % rc)
        self.info = abi.IndexInfo()
        L.bko_info(self._h, C.byref(self.info))

    def close(self):
        if self._h:
            lib().bko_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def entries(self):
        out = []
        for i in range(1, self.info.num_entries + 1):
            e = abi.Entry()
            lib().bko_get_entry(self._h, i, C.byref(e))
            out.append(e)
        return out

    def seq(self):
        n = self.info.concat_len
        return np.ctypeslib.as_array(C.cast(lib().bko_seq(self._h), C.POINTER(C.c_uint8)), shape=(n,))

    def sa_bytes(self):
        n = self.info.concat_len * self.info.sfx_el_size
        return np.ctypeslib.as_array(C.cast(lib().bko_sa(self._h), C.POINTER(C.c_uint8)), shape=(n,))

    def default_params(self, pmode=0, **kw):
        p = abi.AlignParams()
        lib().bko_default_params(self._h, pmode, C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def locate_first_exact(self, probe):
        probe = np.ascontiguousarray(probe, dtype=np.uint8)
        return lib().bko_locate_first_exact(self._h, probe.ctypes.data, len(probe), 0, self.info.concat_len - 1)

    def locate_last_exact(self, probe, lo=0):
        probe = np.ascontiguousarray(probe, dtype=np.uint8)
        return lib().bko_locate_last_exact(self._h, probe.ctypes.data, len(probe), lo, self.info.concat_len - 1)

    def align(self, params, bases, offsets, nthreads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=abi.RESULT_DTYPE)
        st = abi.AlignStats()
        rc = lib().bko_align_batch(self._h, C.byref(params), bases.ctypes.data, offsets.ctypes.data, n,
                                   out.ctypes.data, C.byref(st), nthreads)
        if rc < 0:
            raise RuntimeError("bko_align_batch failed: %d" % rc)
        return out, st

    def align_multi(self, params, bases, offsets, nthreads=1):
        """-r5: (records, loci[n, max_ml_matches], stats)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=abi.RESULT_DTYPE)
        multi = np.zeros((n, params.max_ml_matches), dtype=abi.MULTI_DTYPE)
        st = abi.AlignStats()
        rc = lib().bko_align_batch_multi(self._h, C.byref(params), bases.ctypes.data, offsets.ctypes.data, n,
                                         out.ctypes.data, multi.ctypes.data, C.byref(st), nthreads)
        if rc < 0:
            raise RuntimeError("bko_align_batch_multi failed: %d" % rc)
        return out, multi, st

    def pair(self, params, pe, results, bases=None, offsets=None, len_dist=None, keep=None):
        """keep: uint8 per entry id (index 0 unused), chromosomes that pass the -Z / -z filters inside the pairing."""
        n_pairs = len(results) // 2
        st = abi.PEStats()
        b = np.ascontiguousarray(bases, dtype=np.uint8).ctypes.data if bases is not None else None
        o = np.ascontiguousarray(offsets, dtype=np.uint64).ctypes.data if offsets is not None else None
        ld = len_dist.ctypes.data if len_dist is not None else None
        kp = np.ascontiguousarray(keep, dtype=np.uint8) if keep is not None else None
        rc = lib().bko_pair_reads_filtered(self._h, C.byref(params), C.byref(pe), results.ctypes.data, n_pairs, b, o,
                                           C.byref(st), ld, kp.ctypes.data if kp is not None else None)
        if rc < 0:
            raise RuntimeError("bko_pair_reads failed: %d" % rc)
        return st


# ---- small host helpers shared by tests / bench (not product code) --------------------------------
_A2C = np.full(256, 4, dtype=np.uint8)
for _i, _ch in enumerate(b"ACGT"):
    _A2C[_ch] = _i
    _A2C[_ch + 32] = _i


def read_fasta_reads(path):
    """Tiny FASTA/FASTQ reader for tests: returns (names, bases uint8 codes, offsets uint64)."""
    op = gzip.open if str(path).endswith(".gz") else open
    names, seqs = [], []
    with op(path, "rb") as f:
        data = f.read().split(b"\n")
    i = 0
    while i < len(data):
        ln = data[i]
        if ln.startswith(b">"):
            names.append(ln[1:].split()[0].decode())
            i += 1
            s = b""
            while i < len(data) and not data[i].startswith(b">"):
                s += data[i].strip()
                i += 1
            seqs.append(s)
        elif ln.startswith(b"@"):
            names.append(ln[1:].split()[0].decode())
            seqs.append(data[i + 1].strip())
            i += 4
        else:
            i += 1
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = _A2C[np.frombuffer(b"".join(seqs), dtype=np.uint8)]
    return names, bases, offsets


def pack_reads(reads):
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = np.concatenate(reads).astype(np.uint8) if len(reads) else np.zeros(0, np.uint8)
    return bases, offsets
