#!/usr/bin/env python
"""bench.py -- aligned reads/s of the `biokanga align` hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl bkx|reference]

Workload (config.workload): BASELINE.json configs[1] -- a 3.1 Gbp human-sized synthetic genome with
injected diverged repeats, 20 M simulated 150 bp SE reads carrying 0..4 substitutions, aligned with
`-s3` (MaxTotMM 5 at 150 bp).  Genome, suffix array (GPU prefix doubling)
and reads are generated on the device with fixed seeds; there are no datasets to download.

A "step" is one pass of the align hot path over the rank's 20 M reads.
  value     reads/s with reads already resident in HBM, one byte per base plus the 2-bit copy the fast kernel takes
            (device-pointer C-ABI entry bkx_align_reads_device_packed2, CUDA events); the one-byte-only entry
            is timed beside it (value_byte_layout)
  e2e       the same metric through the compact host call bkx_align_reads_packed2() with pinned HOST buffers: H2D
            of the reads (2 bits per base) and D2H of the 16-byte records are inside the timed region; the 4-bit
            and one-byte-per-base host calls are timed beside it
  roofline  algorithmic bytes of one launch / measured kernel time / measured HBM peak
  cpu_baseline  the CPU restatement (oracle/, "port") on a bounded sample with all host cores
N > 1 (torchrun): rank 0 builds the index and broadcasts it over NCCL/NVLink; every rank aligns its
own 20 M reads (weak scaling); the stats vector is all-reduced each step; time = max over ranks.

--impl reference: the UNMODIFIED reference binary (oracle/_ref/biokanga, built from /root/reference
by oracle/build_ref.sh) run with all host threads on a bounded sample of the same workload; align-phase
time from its own log timestamps.  Falls back to the oracle port when the binary is absent.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


KERNELS_LANE = "align_fast_kernel + align_reads_kernel (deferred reads)"
KERNELS_WAVE = ("one search launch = wave_step_kernel + wave_probe_kernel per search phase (bkx_wave.cuh), then align_fast_kernel "
                "and align_reads_kernel on the reads handed on")


def search_kernels(launches_per_step):
    """Which kernels one device-resident search call launched (the wave path starts ~20 of them)."""
    return KERNELS_WAVE if launches_per_step > 4 else KERNELS_LANE


def ncu_traffic(args, launches_per_step=2):
    """dram__bytes_read.sum + dram__bytes_write.sum of one search launch, from the committed ncu capture -- only when this
    run is the workload AND the kernels that capture was taken on (else null)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_final_traffic.json")))
        if (launches_per_step > 4) != bool(t.get("wave_path", False)):
            return None
        w = t["workload"]
        same = (float(args.genome_mbp) == w["genome_mbp"] and args.reads == w["reads"] and args.read_len == w["read_len"] and
                args.read_subs == w["read_subs"] and args.max_subs == w["max_subs"] and args.seed == w["seed"] and
                args.workload == "se" and args.prefix_k == 0)
        return int(t["dram_bytes_per_launch"]) if same else None
    except Exception:
        return None


def build_workload(args, rank, world, dev, torch, dist):
    """Genome + SA on the device (rank 0 builds, others receive over NCCL), then this rank's reads."""
    from biokanga_b200 import lib as bkx
    from biokanga_b200 import workload as wl
    lens = wl.chrom_layout(int(args.genome_mbp * 1e6))
    ents, n = wl.entries_for(lens)
    t0 = time.time()
    if rank == 0:
        d_seq, ents = wl.make_genome(lens, seed=args.seed, device=dev)
        torch.cuda.synchronize()
        t1 = time.time()
        d_sa = torch.empty(n, dtype=torch.int32, device=dev)
        bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), torch.cuda.current_device())
        torch.cuda.synchronize()
        log("[bench] genome %.2f Gsym in %.1fs, suffix array in %.1fs" % (n / 1e9, t1 - t0, time.time() - t1))
    else:
        d_seq = torch.empty(n, dtype=torch.uint8, device=dev)
        d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    if world > 1:  # index replication GPU0 -> all over NVLink (NCCL broadcast)
        tb = time.time()
        dist.broadcast(d_seq, 0)
        dist.broadcast(d_sa, 0)
        torch.cuda.synchronize()
        if rank == 0:
            log("[bench] index broadcast to %d GPUs in %.2fs" % (world, time.time() - tb))
    if getattr(args, "workload", "se") == "pe":
        d_bases, d_offs = wl.sim_pairs(d_seq, ents, args.reads // 2, args.read_len, seed=args.seed + 100 + rank,
                                       subs=tuple(range(0, args.read_subs + 1)), device=dev)
    else:
        d_bases, d_offs = wl.sim_reads(d_seq, ents, args.reads, args.read_len, seed=args.seed + 100 + rank,
                                       subs=tuple(range(0, args.read_subs + 1)), device=dev)
    torch.cuda.synchronize()
    return d_seq, d_sa, ents, n, d_bases, d_offs


def run_bkx(args):
    import torch
    import torch.distributed as dist
    from biokanga_b200 import abi
    from biokanga_b200 import lib as bkx
    from biokanga_b200 import workload as wl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the bkx path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    d_seq, d_sa, ents, n, d_bases, d_offs = build_workload(args, rank, world, dev, torch, dist)
    t0 = time.time()
    idx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="synth%dM" % args.genome_mbp,
                                device=local, prefix_k=args.prefix_k)
    torch.cuda.synchronize()
    log("[bench] rank %d index resident: %.1f GB HBM, prefix k=%d, %.1fs" % (rank, idx.info.device_bytes / 1e9,
                                                                          idx.info.prefix_k, time.time() - t0))
    host_seq = host_sa = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host_seq = d_seq.cpu().numpy()
        host_sa = d_sa.cpu().numpy().view(np.uint32)
    del d_seq, d_sa
    torch.cuda.empty_cache()

    p = idx.default_params(0, max_subs=args.max_subs)
    nreads = args.reads
    d_out = torch.empty(nreads * 32, dtype=torch.uint8, device=dev)
    d_stats = torch.zeros(C.sizeof(abi.AlignStats) // 8, dtype=torch.int64, device=dev)
    # a dedicated (non-default) stream: launches, events and NCCL all ride on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    pe_mode = args.workload == "pe"
    pe = abi.PEParams()
    pe.pe_proc, pe.pair_min_len, pe.pair_max_len = args.pe_mode, args.pe_min, args.pe_max
    d_pst = torch.zeros(C.sizeof(abi.PEStats) // 8, dtype=torch.int64, device=dev)
    d_ld = torch.zeros(100001, dtype=torch.int32, device=dev)

    # the reads are resident in both layouts: one byte per base (the reference's in-memory layout; read by the general
    # kernel, the orphan recovery and for reads holding an N) and 2 bits per base (what the fast kernel takes)
    d_pk2, d_rflags = wl.pack2_device(d_bases, d_offs)
    torch.cuda.synchronize()
    byte_layout = [False]

    def step():
        d_stats.zero_()
        if byte_layout[0]:
            idx.align_device(p, d_bases.data_ptr(), d_offs.data_ptr(), nreads, args.read_len, d_out.data_ptr(),
                             d_stats.data_ptr(), stream)
        else:
            idx.align_device_packed2(p, d_bases.data_ptr(), d_pk2.data_ptr(), d_rflags.data_ptr(), d_offs.data_ptr(), nreads,
                                     args.read_len, d_out.data_ptr(), d_stats.data_ptr(), stream)
        if pe_mode:
            d_pst.zero_()
            idx.pair_device(p, pe, d_out.data_ptr(), nreads // 2, d_bases.data_ptr(), d_offs.data_ptr(), args.read_len,
                            d_pst.data_ptr(), d_ld.data_ptr(), stream)
        if world > 1:
            dist.all_reduce(d_stats)  # the only collective of the path: global stats counters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = idx.kernel_launches()
    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = (idx.kernel_launches() - l0) // max(1, args.warmup) if args.warmup else 2
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kern_ms = []
    barrier()
    torch.cuda.profiler.start()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
        if args.kernel_times:
            kern_ms.append(idx.last_kernel_ms())  # syncs on the kernel's own events
    barrier()
    torch.cuda.profiler.stop()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    if not kern_ms:
        # one extra, untimed launch gives the kernel's own event-measured duration
        step()
        torch.cuda.synchronize()
        kern_ms = [idx.last_kernel_ms()]
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = world * nreads / (ms_per_step / 1e3)

    res = d_out.cpu().numpy().view(abi.RESULT_DTYPE).copy()
    stats = d_stats.cpu().numpy()
    # the one-byte-per-base entry point on the same reads (what round 1 reported as `value`)
    byte_layout[0] = True
    step()
    barrier()
    evb = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    nb_steps = max(1, min(args.steps, 3))
    evb[0].record()
    for _ in range(nb_steps):
        step()
    evb[1].record()
    barrier()
    tb = torch.tensor([evb[0].elapsed_time(evb[1]) / nb_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    value_bytes = world * nreads / (float(tb.item()) / 1e3)
    same_layouts = bool(d_out.cpu().numpy().view(abi.RESULT_DTYPE).tobytes() == res.tobytes())
    byte_layout[0] = False
    alg_bytes = wl.algorithmic_bytes(res, n, 4, args.read_len)
    kms = float(np.mean(kern_ms))
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kms / 1e3) / 1e9
    nar = np.bincount(res["nar"], minlength=abi.NAR_COUNT)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region.  Headline: the 4-bit packed host call
    # (bkx_align_reads_packed4, the format a loader packs while parsing -- what bkx-align does); the one-byte-per-base
    # call (the reference's in-memory read layout) is timed beside it.
    h_bases = torch.empty(nreads * args.read_len, dtype=torch.uint8).pin_memory()
    h_bases.copy_(d_bases)
    h_packed = torch.from_numpy(bkx.pack_bases4(h_bases.numpy())).pin_memory()
    pk2_np, exc_pos_np, exc_code_np = bkx.pack_bases2(h_bases.numpy())
    h_pk2 = torch.from_numpy(pk2_np).pin_memory()
    h_exc_pos = torch.from_numpy(exc_pos_np.view(np.int64) if len(exc_pos_np) else np.zeros(1, dtype=np.int64)).pin_memory()
    h_exc_code = torch.from_numpy(exc_code_np if len(exc_code_np) else np.zeros(1, dtype=np.uint8)).pin_memory()
    n_exc = int(len(exc_pos_np))
    h_offs = (torch.arange(nreads + 1, dtype=torch.int64) * args.read_len).pin_memory()
    h_out = torch.empty(nreads * 32, dtype=torch.uint8).pin_memory()
    h_out16 = torch.empty(nreads * 16, dtype=torch.uint8).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    hst = abi.AlignStats()
    h_pst = abi.PEStats()

    def e2e_call(form):
        if form == "packed2":   # the compact host interface: 2 bits per base in, 16-byte records out
            idx.align_packed2_ptr(p, h_pk2.data_ptr(), None, args.read_len, h_exc_pos.data_ptr() if n_exc else None,
                                  h_exc_code.data_ptr() if n_exc else None, n_exc, nreads, h_out16.data_ptr(), hst,
                                  pe if pe_mode else None, h_pst if pe_mode else None, None)
        elif pe_mode:  # align + pair fused: every slice is paired (orphans recovered) while it is still on the GPU
            idx.align_pairs_ptr(p, pe, (h_packed if form == "packed4" else h_bases).data_ptr(), h_offs.data_ptr(), nreads // 2,
                                h_out.data_ptr(), hst, h_pst, None, packed=form == "packed4")
        elif form == "packed4":
            idx.align_packed4_ptr(p, h_packed.data_ptr(), h_offs.data_ptr(), nreads, h_out.data_ptr(), hst)
        else:
            idx.align_ptr(p, h_bases.data_ptr(), h_offs.data_ptr(), nreads, h_out.data_ptr(), hst)

    def time_e2e(form):
        e2e_call(form)  # warm-up
        barrier()
        reps = []
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            e2e_call(form)
            reps.append(round((time.perf_counter() - t1) * 1e3, 2))
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        log("[bench] e2e %s: per-call ms %s, kernel ms %.2f" % (form, reps, idx.last_kernel_ms()))
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if form == "packed2":
            got = bkx.expand_results16(h_out16.numpy().view(abi.RESULT16_DTYPE), fixed_len=args.read_len)
        else:
            got = h_out.numpy().view(abi.RESULT_DTYPE)
        ok = all(bool(np.array_equal(got[f], res[f])) for f in ("nar", "hit_rslt", "strand", "chrom_id", "match_loci", "match_len",
                                                                "mismatches", "low_mm", "nxt_low_mm", "low_hit_instances", "flags"))
        return world * nreads / float(te.item()), ok

    e2e_value, same = time_e2e("packed2")
    e2e_p4_value, same_4 = time_e2e("packed4")
    e2e_bytes_value, same_b = time_e2e("bytes")
    same = same and same_4 and same_b

    line = {
        "metric": "aligned reads/sec (150bp, <=4 subs)", "value": value, "unit": "reads/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": ("configs[2] shape: %.1f Gbp synthetic genome + injected repeats, %d x 2x%d bp PE pairs per GPU, "
                                "0..%d subs, -s%d -U%d -d%d -D%d" % (args.genome_mbp / 1e3, nreads // 2, args.read_len,
                                                                        args.read_subs, args.max_subs, args.pe_mode, args.pe_min,
                                                                        args.pe_max)) if pe_mode else
                               ("%s: %s synthetic genome + injected repeats, %d x %d bp SE reads per GPU, "
                                "0..%d subs, -s%d" % ("configs[0]" if args.genome_mbp < 100 else "configs[1]",
                                                       ("%.0f Mbp" % args.genome_mbp) if args.genome_mbp < 100 else ("%.1f Gbp" % (args.genome_mbp / 1e3)),
                                                       nreads, args.read_len, args.read_subs, args.max_subs)),
                   "genome_symbols": int(n), "reads_per_gpu": nreads, "read_len": args.read_len,
                   "max_subs_per_100bp": args.max_subs, "prefix_k": int(idx.info.prefix_k),
                   "l2_policy": "inputs larger than L2 (index %.1f GB, reads %.1f GB per step)" % (
                       idx.info.device_bytes / 1e9, nreads * args.read_len / 1e9),
                   "resident_layout": "reads in HBM one byte per base + their 2-bit copy (bkx_align_reads_device_packed2)",
                   "parallelism": "reads sharded over %d GPU(s), index replicated" % world},
        "value_byte_layout": {"value": value_bytes, "call": "bkx_align_reads_device (reads resident one byte per base only)",
                              "same_records": same_layouts},
        "e2e": {"value": e2e_value, "unit": "reads/s",
                "h2d_bytes_per_step": int((nreads * args.read_len + 3) // 4 + n_exc * 9),
                "d2h_bytes_per_step": int(nreads * 16), "matches_device_run": same,
                "call": ("bkx_align_pairs_packed2" if pe_mode else "bkx_align_reads_packed2") +
                        " (pinned host buffers: reads 2 bits per base + list of the non-ACGT bases, fixed read length, "
                        "16-byte records back)",
                "bytes_per_read_over_pcie": ((nreads * args.read_len + 3) // 4 + n_exc * 9 + nreads * 16) / nreads,
                "packed4_call": {"value": e2e_p4_value, "call": "bkx_align_pairs_packed4" if pe_mode else "bkx_align_reads_packed4",
                                 "h2d_bytes_per_step": int((nreads * args.read_len + 1) // 2 + (nreads + 1) * 8),
                                 "d2h_bytes_per_step": int(nreads * 32)},
                "byte_per_base_call": {"value": e2e_bytes_value, "call": "bkx_align_pairs" if pe_mode else "bkx_align_reads",
                                       "h2d_bytes_per_step": int(nreads * args.read_len + (nreads + 1) * 8),
                                       "d2h_bytes_per_step": int(nreads * 32)}},
        "gpu_launches": int(args.steps * launches_per_step),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args, launches_per_step), "kernel": search_kernels(launches_per_step),
                     "kernel_ms": kms, "algorithmic_bytes_per_launch": int(alg_bytes), "bytes_per_read": alg_bytes / nreads,
                     "peak_source": peak_src,
                     "traffic_source": "profiles/r02_final_traffic.json (ncu, one search launch of the same kernels; null when "
                                       "this run used other kernels or another workload)"},
        "clocks": clocks,
        "classes": {abi.NAR_CODES[i]: int(nar[i]) for i in range(abi.NAR_COUNT) if nar[i]},
        "stats_reads_all_ranks": int(stats[-1]),
    }
    if pe_mode:
        pst = d_pst.cpu().numpy()
        line["pe"] = {"pairs_per_gpu": nreads // 2, "accepted_pairs": int(pst[1]), "recovered_orphans": int(pst[3]),
                      "unaligned_pairs": int(pst[0]), "pairs_per_s": value / 2}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not pe_mode:   # reported once: rank 0 of the single-GPU run
        line["cpu_baseline"] = cpu_baseline_port(args, host_seq, host_sa, ents, h_bases.numpy(), res)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE.json configs[4]: substitutions 0..8 x read length 50..300 on the configs[1] genome, at 1/2/4/8 GPUs (under
    torchrun the index is broadcast over NCCL, every rank aligns its own reads of every cell, a cell's time is the maximum
    over the ranks).  One table row per (L, -s): device-resident reads/s (all ranks), class histogram, algorithmic-bytes
    roofline and -- on rank 0 -- a parity check of a sample of the records against the CPU oracle, whose throughput on
    that sample is the host-CPU figure of the row.  Writes gpurun_out/sweep[_Ngpu].json; prints one summary JSON line."""
    import torch
    import torch.distributed as dist
    from biokanga_b200 import abi
    from biokanga_b200 import lib as bkx
    from biokanga_b200 import workload as wl
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args.reads = 1024   # the reads build_workload makes are not used: every cell draws its own
    d_seq, d_sa, ents, n, _, _ = build_workload(args, rank, world, dev, torch, dist)
    idx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="sweep", device=local, prefix_k=args.prefix_k)
    oidx = None
    if rank == 0:
        oidx = po.OracleIndex(seq=d_seq.cpu().numpy(), sa=d_sa.cpu().numpy().view(np.uint32), el_size=4, entries=ents)
    del d_sa
    torch.cuda.empty_cache()
    peak, _ = measured_peak()
    nreads = args.sweep_reads
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    rows = []
    all_ok = True
    cores = os.cpu_count() or 1
    lens_ = [int(v) for v in args.sweep_lens.split(",")]
    subs_ = [int(v) for v in args.sweep_subs.split(",")]
    for L in lens_:
        for s_ in subs_:
            for mmd in ((1, 2) if s_ in (3, 8) else (1,)):
                max_tot = 0 if s_ == 0 else max(1, (L * s_ + 50) // 100)
                d_bases, d_offs = wl.sim_reads(d_seq, ents, nreads, L, seed=args.seed + 7 * L + s_ + 1000 * rank,
                                               subs=tuple(range(0, max_tot + 2)), device=dev)
                d_pk2, d_rflags = wl.pack2_device(d_bases, d_offs)
                p = idx.default_params(0, max_subs=s_, min_edit_dist=mmd)
                d_out = torch.empty(nreads * 32, dtype=torch.uint8, device=dev)
                ms = []
                for rep in range(3):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    idx.align_device_packed2(p, d_bases.data_ptr(), d_pk2.data_ptr(), d_rflags.data_ptr(), d_offs.data_ptr(), nreads, L,
                                             d_out.data_ptr(), None, tstream.cuda_stream)
                    torch.cuda.synchronize()
                    t = torch.tensor([idx.last_kernel_ms()], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms.append(float(t.item()))
                if rank != 0:
                    del d_bases, d_offs, d_out, d_pk2, d_rflags
                    continue
                res = d_out.cpu().numpy().view(abi.RESULT_DTYPE)
                m = min(args.sweep_check, nreads)
                bases = d_bases[:m * L].cpu().numpy()
                offs = np.arange(m + 1, dtype=np.uint64) * L
                tq = time.perf_counter()
                exp, _ = oidx.align(oidx.default_params(0, max_subs=s_, min_edit_dist=mmd), bases, offs, nthreads=cores)
                cpu_s = time.perf_counter() - tq
                ok = all(np.array_equal(res[f][:m], exp[f]) for f in abi.RESULT_DTYPE.names)
                all_ok &= ok
                nar = np.bincount(res["nar"], minlength=abi.NAR_COUNT)
                best = min(ms[1:])
                ab = wl.algorithmic_bytes(res, n, 4, L)
                rows.append({"L": L, "s": s_, "e": mmd, "max_tot_mm": max_tot, "reads_per_gpu": nreads, "n_gpus": world, "ms": best,
                             "reads_per_s": world * nreads / (best / 1e3), "roofline_frac": ab / (best / 1e3) / 1e9 / peak,
                             "bytes_per_read": ab / nreads, "parity_sample": m, "parity_ok": bool(ok),
                             "host_cpu_port_reads_per_s": m / cpu_s, "host_cores": cores,
                             "classes": {abi.NAR_CODES[i]: int(nar[i]) for i in range(abi.NAR_COUNT) if nar[i]}})
                log("[sweep] L=%d -s%d -e%d: %.1f M reads/s on %d GPU(s), frac %.2f per GPU, parity %s, host CPU %.0f k reads/s, %s" % (
                    L, s_, mmd, rows[-1]["reads_per_s"] / 1e6, world, rows[-1]["roofline_frac"], ok, m / cpu_s / 1e3, rows[-1]["classes"]))
                del d_bases, d_offs, d_out, d_pk2, d_rflags
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "sweep.json" if world == 1 else "sweep_%dgpu.json" % world
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
        print(json.dumps({"sweep_rows": len(rows), "n_gpus": world, "all_parity_ok": bool(all_ok),
                          "min_reads_per_s": min(r["reads_per_s"] for r in rows), "max_reads_per_s": max(r["reads_per_s"] for r in rows)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_hexaploid(args):
    """BASELINE.json configs[3]: wheat-scale hexaploid (3 sub-genomes x 7 chromosomes, >= 4e9 symbols => 5-byte suffix
    elements), 150 bp SE reads, one GPU.  The suffix array is built on the device in planes (the reference's own
    `biokanga index` needs ~6 bytes/symbol of host RAM and hours at this size); everything after that is the same
    C-ABI path as configs[1].  Prints one JSON line; writes gpurun_out/hexaploid.json."""
    import torch
    from biokanga_b200 import abi
    from biokanga_b200 import lib as bkx
    from biokanga_b200 import workload as wl
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the bkx path has no CPU fallback")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    chrom_len = int(args.genome_mbp * 1e6) // 21
    t0 = time.time()
    d_seq, ents = wl.make_hexaploid(chrom_len, seed=args.seed, device=dev)
    n = int(d_seq.numel())
    torch.cuda.synchronize()
    t1 = time.time()
    five = n >= 4_000_000_000
    if five:   # 5-byte elements back to back, as in the .sfx file: the layout the search reads (one fetch per element)
        d_sa = torch.zeros(n * 5 + 16, dtype=torch.uint8, device=dev)
        torch.cuda.empty_cache()
        bkx.build_suffix_array_packed5(d_seq.data_ptr(), n, d_sa.data_ptr(), 0, 0)
    else:
        d_sa = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.empty_cache()
        bkx.build_suffix_array_planes(d_seq.data_ptr(), n, d_sa.data_ptr(), None, 0, 0)
    torch.cuda.synchronize()
    t2 = time.time()
    log("[bench] hexaploid genome %.2f Gsym in %.1fs, suffix array (%d-byte elements) in %.1fs" % (
        n / 1e9, t1 - t0, 5 if five else 4, t2 - t1))
    # the reads are drawn before the index takes the rest of the HBM (k = 17 table: 69 GB beside the 70 GB suffix array)
    nreads, L = args.reads, args.read_len
    d_bases, d_offs = wl.sim_reads(d_seq, ents, nreads, L, seed=args.seed + 100, subs=tuple(range(0, args.read_subs + 1)),
                                   device=dev)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if five:
        idx = bkx.Index.from_packed5(d_seq.data_ptr(), n, d_sa.data_ptr(), ents, name="hexaploid%dM" % args.genome_mbp, device=0,
                                     prefix_k=args.prefix_k)
    else:
        idx = bkx.Index.from_planes(d_seq.data_ptr(), n, d_sa.data_ptr(), None, ents, name="hexaploid%dM" % args.genome_mbp,
                                    device=0, prefix_k=args.prefix_k)
    torch.cuda.synchronize()
    t3 = time.time()
    log("[bench] index resident: %.1f GB HBM, prefix k=%d, %.1fs" % (idx.info.device_bytes / 1e9, idx.info.prefix_k, t3 - t2))
    el = int(idx.info.sfx_el_size)
    p = idx.default_params(0, max_subs=args.max_subs)
    d_out = torch.empty(nreads * 32, dtype=torch.uint8, device=dev)
    d_stats = torch.zeros(C.sizeof(abi.AlignStats) // 8, dtype=torch.int64, device=dev)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)

    d_pk2, d_rflags = wl.pack2_device(d_bases, d_offs)   # the reads are resident one byte per base and 2-bit packed
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    def step():
        d_stats.zero_()
        idx.align_device_packed2(p, d_bases.data_ptr(), d_pk2.data_ptr(), d_rflags.data_ptr(), d_offs.data_ptr(), nreads, L,
                                 d_out.data_ptr(), d_stats.data_ptr(), tstream.cuda_stream)

    l0 = idx.kernel_launches()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = (idx.kernel_launches() - l0) // max(1, args.warmup) if args.warmup else 2
    sampler = ClockSampler(0)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    step()
    torch.cuda.synchronize()
    kms = idx.last_kernel_ms()
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    res = d_out.cpu().numpy().view(abi.RESULT_DTYPE)
    stats = d_stats.cpu().numpy()
    alg = wl.algorithmic_bytes(res, n, el, L)
    peak, peak_src = measured_peak()
    nar = np.bincount(res["nar"], minlength=abi.NAR_COUNT)

    # e2e through the host-buffer call
    h_bases = torch.empty(nreads * L, dtype=torch.uint8).pin_memory()
    h_bases.copy_(d_bases)
    pk2_np, exc_pos_np, exc_code_np = bkx.pack_bases2(h_bases.numpy())
    h_pk2 = torch.from_numpy(pk2_np).pin_memory()
    n_exc = int(len(exc_pos_np))
    h_exc_pos = torch.from_numpy(exc_pos_np.view(np.int64) if n_exc else np.zeros(1, dtype=np.int64)).pin_memory()
    h_exc_code = torch.from_numpy(exc_code_np if n_exc else np.zeros(1, dtype=np.uint8)).pin_memory()
    h_out16 = torch.empty(nreads * 16, dtype=torch.uint8).pin_memory()
    hst = abi.AlignStats()

    def e2e_call():
        idx.align_packed2_ptr(p, h_pk2.data_ptr(), None, L, h_exc_pos.data_ptr() if n_exc else None,
                              h_exc_code.data_ptr() if n_exc else None, n_exc, nreads, h_out16.data_ptr(), hst)
    e2e_call()
    e2e_steps = max(1, min(args.steps, 3))
    tq = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - tq) / e2e_steps
    got = bkx.expand_results16(h_out16.numpy().view(abi.RESULT16_DTYPE), fixed_len=L)
    same = all(bool(np.array_equal(got[f], res[f])) for f in ("nar", "hit_rslt", "strand", "chrom_id", "match_loci", "match_len",
                                                              "mismatches", "low_mm", "nxt_low_mm", "low_hit_instances", "flags"))

    # size-independent check: every accepted record carries exactly the mismatches found at its locus
    acc = np.nonzero(res["nar"] == abi.NAR_ACCEPTED)[0]
    pick = acc[:: max(1, len(acc) // 200000)]
    ofs_of = {int(e["entry_id"]): int(e["start_ofs"]) for e in ents}
    g0 = torch.from_numpy(np.array([ofs_of[int(c)] for c in res["chrom_id"][pick]], dtype=np.int64) +
                          res["match_loci"][pick].astype(np.int64)).to(dev)
    ar = torch.arange(L, device=dev)
    gw = d_seq[g0[:, None] + ar[None, :]]
    rd = d_bases.view(-1, L)[torch.from_numpy(pick.astype(np.int64)).to(dev)]
    minus = torch.from_numpy((res["strand"][pick] == ord("-"))).to(dev)
    rrc = torch.flip(rd, dims=[1])
    rrc = torch.where(rrc < 4, 3 - rrc, rrc)
    rd = torch.where(minus[:, None], rrc, rd)
    mm = (rd != gw).sum(dim=1).cpu().numpy()
    loci_ok = bool(np.array_equal(mm, res["mismatches"][pick]))

    line = {
        "metric": "aligned reads/sec (150bp, <=4 subs)", "value": nreads / (ms_per_step / 1e3), "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[3]: %.1f Gbp synthetic hexaploid (B, D = A + 2-5 %% substitutions + block shuffles), "
                               "%d-byte suffix elements, %d x %d bp SE reads per step (x %d steps = %d reads), 0..%d subs, -s%d" % (
                                   args.genome_mbp / 1e3, el, nreads, L, args.steps, nreads * args.steps, args.read_subs,
                                   args.max_subs),
                   "genome_symbols": n, "reads_per_step": nreads, "read_len": L, "prefix_k": int(idx.info.prefix_k),
                   "index_gb": idx.info.device_bytes / 1e9 + n * el / 1e9,
                   "build_s": {"genome": t1 - t0, "suffix_array_gpu": t2 - t1, "index_tables": t3 - t2},
                   "l2_policy": "inputs larger than L2"},
        "e2e": {"value": nreads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": int((nreads * L + 3) // 4 + n_exc * 9),
                "d2h_bytes_per_step": int(nreads * 16), "matches_device_run": same,
                "call": "bkx_align_reads_packed2 (pinned host buffers, 2 bits per base in, 16-byte records out)"},
        "gpu_launches": int(args.steps * launches_per_step),
        "roofline": {"bound": "hbm", "achieved": alg / (kms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (kms / 1e3) / 1e9 / peak, "traffic": None,
                     "kernel": search_kernels(launches_per_step), "kernel_ms": kms,
                     "algorithmic_bytes_per_launch": int(alg), "bytes_per_read": alg / nreads, "peak_source": peak_src},
        "clocks": clocks,
        "classes": {abi.NAR_CODES[i]: int(nar[i]) for i in range(abi.NAR_COUNT) if nar[i]},
        "accepted_loci_verified": {"records": int(len(pick)), "all_mismatch_counts_exact": loci_ok},
        "stats_reads": int(stats[-1]),
    }
    # oracle on a sample, with the whole index copied to host memory (84 GB at 14 G symbols)
    if not args.no_cpu_baseline:
        import psutil
        need = n * (1 + el) + (8 << 30)
        if psutil.virtual_memory().available < need:
            line["cpu_baseline"] = {"skipped": "host RAM: %.0f GB needed" % (need / 1e9)}
        else:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyoracle as po
            tc = time.time()
            host_seq = d_seq.cpu().numpy()
            if el == 4:
                host_sa = d_sa.cpu().numpy().view(np.uint32)
            else:
                host_sa = d_sa[:n * 5].cpu().numpy()
            log("[bench] index copied to host in %.1fs" % (time.time() - tc))
            oidx = po.OracleIndex(seq=host_seq, sa=host_sa, el_size=el, entries=ents)
            m = min(args.cpu_sample, 300000, nreads)   # the 14 G-symbol search is several times slower per read
            cores = os.cpu_count() or 1
            hb = h_bases.numpy()[:m * L]
            ho = np.arange(m + 1, dtype=np.uint64) * L
            tq = time.perf_counter()
            exp, _ = oidx.align(oidx.default_params(0, max_subs=args.max_subs), hb, ho, nthreads=cores)
            dt = time.perf_counter() - tq
            ok = all(np.array_equal(exp[f], res[f][:m]) for f in abi.RESULT_DTYPE.names)
            line["cpu_baseline"] = {"value": m / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": "first %d of the %d reads, %.1fs" % (m, nreads, dt), "parity_with_gpu": bool(ok)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", "hexaploid.json"), "w"), indent=1)
    print(json.dumps(line), flush=True)


def cpu_baseline_port(args, host_seq, host_sa, ents, h_bases, gpu_res):
    """The CPU restatement on a bounded sample of the same reads, all host cores; also re-checks parity."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    cores = os.cpu_count() or 1
    oidx = po.OracleIndex(seq=host_seq, sa=host_sa, el_size=4, entries=ents)
    p = oidx.default_params(0, max_subs=args.max_subs)
    m = min(args.cpu_sample, args.reads)
    L = args.read_len
    bases = h_bases[:m * L]
    offs = (np.arange(m + 1, dtype=np.uint64) * L)
    t0 = time.perf_counter()
    res, st = oidx.align(p, bases, offs, nthreads=cores)
    dt = time.perf_counter() - t0
    ok = all(np.array_equal(res[f], gpu_res[f][:m]) for f in ("nar", "strand", "chrom_id", "match_loci", "mismatches",
                                                              "low_hit_instances", "seeds", "cands"))
    return {"value": m / dt, "unit": "reads/s", "cores": cores, "kind": "port",
            "sample": "first %d of the %d reads, %.1fs" % (m, args.reads, dt), "parity_with_gpu": bool(ok)}


def write_fasta_fast(path, reads2d, alphabet):
    """Vectorised FASTA writer: fixed-width names '>r000000001'."""
    n, L = reads2d.shape
    rec = np.empty((n, 12 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    ids = np.arange(1, n + 1, dtype=np.int64)
    for d in range(9):
        rec[:, 10 - d] = ord("0") + (ids // (10 ** d)) % 10
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + L] = alphabet[reads2d]
    rec[:, 12 + L] = ord("\n")
    with open(path, "wb") as f:
        f.write(rec.tobytes())


def run_reference(args):
    """Reference arm: the reference's own CPU implementation on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    from biokanga_b200 import abi
    from biokanga_b200 import lib as bkx
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    import synth
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "needs the GPU box to generate the 3.1 Gbp workload"}))
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    # The reference's main thread sleeps a fixed 5 s after starting its workers and only then joins them
    # (Aligner.cpp:8797-8801): an align phase that its log times at ~5.0 s measured the sleep, not the work.  So every
    # step must keep the workers busy well beyond that: the sample grows with the core count, every step is checked
    # against a 7 s floor, and a step that lands under it doubles the sample (up to the pool generated here) and is redone.
    pool = int(min(48_000_000, max(args.ref_sample, 400_000 * cores)))
    sample = int(min(pool, max(args.ref_sample, 250_000 * cores)))
    floor_s = 7.0
    saved_reads = args.reads
    args.reads = pool
    d_seq, d_sa, ents, n, d_bases, d_offs = build_workload(args, 0, 1, dev, torch, None)
    host_seq = d_seq.cpu().numpy()
    host_sa = d_sa.cpu().numpy().view(np.uint32)
    h_bases = d_bases.cpu().numpy()
    del d_seq, d_sa, d_bases
    torch.cuda.empty_cache()
    L = args.read_len
    use_bin = os.path.exists(po.REF_BIN) and not args.ref_port
    times = []
    floor_hit = False
    ref_counts = None
    if use_bin:
        tmp = tempfile.mkdtemp(prefix="bkxref", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            t0 = time.time()
            bkx.write_sfx(os.path.join(tmp, "g.sfx"), host_seq, host_sa, 4, ents, name="synth")
            log("[bench] wrote %.1f GB .sfx in %.1fs" % (os.path.getsize(os.path.join(tmp, "g.sfx")) / 1e9, time.time() - t0))
            del host_seq, host_sa

            def one_step(m):
                """One unmodified `biokanga align` over the first m reads of the pool: (align-phase seconds from its own log
                timestamps, reads it says it aligned)."""
                write_fasta_fast(os.path.join(tmp, "r.fa"), h_bases[:m * L].reshape(m, L), synth.BASES)
                subprocess.run([po.REF_BIN, "align", "-I", "g.sfx", "-i", "r.fa", "-s%d" % args.max_subs, "-M0", "-o",
                                "out.csv", "-F", "run.log", "-T%d" % min(cores, 128)], cwd=tmp, check=True,
                               stdout=subprocess.DEVNULL)
                ts, done = {}, None
                for ln in open(os.path.join(tmp, "run.log"), errors="replace"):
                    mt = re.match(r"\[\w+ +\d+ (\d+):(\d+):(\d+)\.(\d+) \d+\]", ln)
                    if not mt:
                        continue
                    t = int(mt.group(1)) * 3600 + int(mt.group(2)) * 60 + int(mt.group(3)) + int(mt.group(4)) / 1000.0
                    if "Now aligning with minimum core size" in ln:
                        ts["a"] = t
                    md = re.search(r"Alignment of (\d+) from (\d+) loaded completed", ln)
                    if md:
                        ts["b"] = t
                        done = (int(md.group(1)), int(md.group(2)))
                return (ts["b"] - ts["a"]) % 86400, done

            s = 0
            while s < total_steps:
                dt, done = one_step(sample)
                log("[bench] reference step %d: %d reads, align phase %.2fs (its log: aligned %s of %s loaded)" % (
                    s, sample, dt, done[0] if done else "?", done[1] if done else "?"))
                if dt < floor_s:
                    if sample < pool:   # the 5 s sleep, not the work: more reads, start over
                        sample = int(min(pool, sample * 2))
                        times, s = [], 0
                        log("[bench] reference step under the %.0f s floor of its fixed sleep: sample raised to %d reads" % (floor_s, sample))
                        continue
                    floor_hit = True
                if s >= args.warmup:
                    times.append(dt)
                    ref_counts = done
                s += 1
        finally:
            subprocess.run(["rm", "-rf", tmp])
        kind = "reference"
    else:
        oidx = po.OracleIndex(seq=host_seq, sa=host_sa, el_size=4, entries=ents)
        p = oidx.default_params(0, max_subs=args.max_subs)
        sample = int(min(sample, args.ref_sample))   # the port has no sleep: the plain bounded sample
        for s in range(total_steps):
            bases = h_bases[:sample * L]
            offs = np.arange(sample + 1, dtype=np.uint64) * L
            t0 = time.perf_counter()
            oidx.align(p, bases, offs, nthreads=cores)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
        kind = "port"
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "aligned reads/sec (150bp, <=4 subs)", "value": value, "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: %.1f Gbp synthetic genome + injected repeats, %d bp SE reads, 0..%d subs, "
                               "-s%d; each step a bounded sample of %d reads" % (args.genome_mbp / 1e3, L,
                                                                                 args.read_subs, args.max_subs, sample),
                   "genome_symbols": int(n), "read_len": L, "max_subs_per_100bp": args.max_subs},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": "%d reads per step (one unmodified `biokanga align -T%d` run per step); align phase from the "
                                   "reference's own log timestamps, every step above the %.0f s floor of its fixed 5 s sleep: %s" % (
                                       sample, min(cores, 128), floor_s, not floor_hit)
                         if kind == "reference" else "%d reads per step, %d threads" % (sample, cores),
                         "step_seconds": [round(t, 2) for t in times], "floor_hit": floor_hit,
                         "reference_log_reads_aligned": ref_counts[0] if ref_counts else None},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.reads = saved_reads
    print(json.dumps(line), flush=True)


def run_dropin(args):
    """Full-size drop-in check (SURVEY.md section 8(c) comparison surface + 8(d) "end-to-end wall"): the reference's
    `biokanga align` and this repo's `bkx-align` front end on the SAME files -- a 3.1 Gbp .sfx and a FASTA of
    --ref-sample reads -- with the same options; compares the -M0 CSV rows and the alignment-summary block of the
    logs, and reports both walls (process start to exit).  Needs the GPU box and oracle/_ref."""
    import hashlib
    import torch
    from biokanga_b200 import lib as bkx
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    import synth
    cli = os.path.join(os.path.dirname(bkx.LIB_PATH), "bkx-align")
    if not (torch.cuda.is_available() and os.path.exists(po.REF_BIN) and os.path.exists(cli)):
        print(json.dumps({"dropin": "unavailable", "why": "needs a GPU, oracle/_ref and biokanga_b200/bkx-align"}))
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    args.reads = sample
    d_seq, d_sa, ents, n, d_bases, d_offs = build_workload(args, 0, 1, dev, torch, None)
    host_seq = d_seq.cpu().numpy()
    host_sa = d_sa.cpu().numpy().view(np.uint32)
    h_bases = d_bases.cpu().numpy()
    del d_seq, d_sa, d_bases, d_offs
    torch.cuda.empty_cache()
    L = args.read_len
    tmp = tempfile.mkdtemp(prefix="bkxdrop", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    def summary(path):
        keep, on = [], False
        for ln in open(path, errors="replace"):
            body = ln.split("](biokanga) ", 1)[1] if "](biokanga) " in ln else ln
            if "Alignment of" in body and "completed" in body:
                on = True
            if on and ("Reporting of aligned result set" in body or "Exit code" in body or "Total processing time" in body):
                continue
            if on:
                keep.append(body.rstrip("\n"))
        return keep

    def digest(path):
        rows = open(path, "rb").read().splitlines()
        rows.sort()
        h = hashlib.md5()
        for r in rows:
            h.update(r)
            h.update(b"\n")
        return len(rows), h.hexdigest()

    try:
        bkx.write_sfx(os.path.join(tmp, "g.sfx"), host_seq, host_sa, 4, ents, name="synth")
        del host_seq, host_sa
        rd = h_bases[:sample * L].reshape(sample, L)
        if args.workload == "pe":  # PE1 / PE2 of a pair are adjacent in the simulated batch
            write_fasta_fast(os.path.join(tmp, "r.fa"), rd[0::2], synth.BASES)
            write_fasta_fast(os.path.join(tmp, "r2.fa"), rd[1::2], synth.BASES)
            common = ["align", "-I", "g.sfx", "-i", "r.fa", "-u", "r2.fa", "-U%d" % args.pe_mode, "-d%d" % args.pe_min,
                      "-D%d" % args.pe_max, "-s%d" % args.max_subs, "-M0"]
        else:
            write_fasta_fast(os.path.join(tmp, "r.fa"), rd, synth.BASES)
            common = ["align", "-I", "g.sfx", "-i", "r.fa", "-s%d" % args.max_subs, "-M0"]
        if args.dropin_front_end_only:   # timing of the front end alone (several runs; the first one warms the page cache)
            walls = []
            for rep in range(3):
                t0 = time.time()
                subprocess.run([cli] + common + ["-o", "bkx.csv", "-F", "bkx.log"], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
                walls.append(round(time.time() - t0, 3))
            log(open(os.path.join(tmp, "bkx.log")).read())
            print(json.dumps({"dropin_front_end_only": walls, "reads": sample, "workload": args.workload}), flush=True)
            return
        t0 = time.time()
        subprocess.run([po.REF_BIN] + common + ["-o", "ref.csv", "-F", "ref.log", "-T%d" % min(cores, 128)], cwd=tmp,
                       check=True, stdout=subprocess.DEVNULL)
        t_ref = time.time() - t0
        t0 = time.time()
        subprocess.run([cli] + common + ["-o", "bkx.csv", "-F", "bkx.log"], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        t_bkx = time.time() - t0
        n_ref, md_ref = digest(os.path.join(tmp, "ref.csv"))
        n_bkx, md_bkx = digest(os.path.join(tmp, "bkx.csv"))
        s_ref, s_bkx = summary(os.path.join(tmp, "ref.log")), summary(os.path.join(tmp, "bkx.log"))
        line = {"dropin": "%s files: %.1f Gbp .sfx (%.1f GB), %d x %d bp reads, %s" % (
                    "configs[2] shape" if args.workload == "pe" else "configs[1]", args.genome_mbp / 1e3,
                    os.path.getsize(os.path.join(tmp, "g.sfx")) / 1e9, sample, L, " ".join(common[6 if args.workload == "pe" else 4:])),
                "csv_rows": [n_ref, n_bkx], "csv_md5_sorted": [md_ref, md_bkx], "csv_identical": md_ref == md_bkx and n_ref == n_bkx,
                "summary_lines": [len(s_ref), len(s_bkx)], "summary_identical": s_ref == s_bkx,
                "wall_s": {"reference": t_ref, "bkx-align": t_bkx, "reference_threads": min(cores, 128)}}
        if s_ref != s_bkx:
            line["summary_diff"] = [x for x in zip(s_ref, s_bkx) if x[0] != x[1]][:6]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        for nm in ("ref.log", "bkx.log"):
            subprocess.run(["cp", os.path.join(tmp, nm), os.path.join(ROOT, "gpurun_out", "dropin_" + nm)])
        json.dump(line, open(os.path.join(ROOT, "gpurun_out", "dropin.json"), "w"), indent=1)
        print(json.dumps(line), flush=True)
    finally:
        subprocess.run(["rm", "-rf", tmp])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bkx", choices=["bkx", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=3100.0)
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--read-subs", type=int, default=4, help="reads carry 0..this many substitutions")
    ap.add_argument("--max-subs", type=int, default=3, help="-s: allowed substitutions per 100 bp")
    ap.add_argument("--prefix-k", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--cpu-sample", type=int, default=4000000,
                    help="reads of the workload the CPU baseline aligns (about 10-30 s of host time on 16 cores at configs[1])")
    ap.add_argument("--ref-sample", type=int, default=4000000)
    ap.add_argument("--ref-port", action="store_true", help="reference arm: use the oracle port even if the binary exists")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="se", choices=["se", "pe", "hexaploid"],
                    help="se: configs[1] (default, the headline); pe: configs[2] shape -- 2x150 bp pairs, pairing + orphan recovery")
    ap.add_argument("--pe-mode", type=int, default=1, help="-U mode for --workload pe (1 = recover orphans)")
    ap.add_argument("--pe-min", type=int, default=200)
    ap.add_argument("--pe-max", type=int, default=1000)
    ap.add_argument("--sweep", action="store_true", help="configs[4]: substitutions 0..8 x read length 50..300 table")
    ap.add_argument("--sweep-reads", type=int, default=2_000_000)
    ap.add_argument("--sweep-check", type=int, default=20000)
    ap.add_argument("--sweep-lens", default="50,75,100,150,200,250,300")
    ap.add_argument("--sweep-subs", default="0,1,2,3,4,5,6,7,8")
    ap.add_argument("--dropin", action="store_true",
                    help="run the reference binary and bkx-align on the same 3.1 Gbp files and compare their outputs")
    ap.add_argument("--dropin-front-end-only", action="store_true", help="--dropin without the reference run: bkx-align walls only")
    ap.add_argument("--kernel-times", action="store_true", help="read the kernel's own events after every step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "bkx":
        log("[bench] note: fewer than 3 warm-up steps requested")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node %d" % args.gpus)
    if args.sweep:
        run_sweep(args)
    elif args.dropin:
        run_dropin(args)
    elif args.workload == "hexaploid":
        run_hexaploid(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_bkx(args)


if __name__ == "__main__":
    main()
