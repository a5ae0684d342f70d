#!/usr/bin/env python
"""A/B timing of the align kernels on ONE resident index: builds the configs[1] workload once (or a smaller one), then
times the device-resident call under several settings of the library's tuning environment variables and checks that
every setting returns the same records.

    python profiles/ab_kernel.py [--genome-mbp 3100] [--reads 20000000] [--read-len 150] [--max-subs 3] \
        --set BKX_SCAN_ITERS=0 --set BKX_SCAN_ITERS=8 --set "BKX_SCAN_ITERS=8 BKX_X=1"

Prints one line per setting: ms per launch (best / median of --reps), reads/s, identical-to-first."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mbp", type=float, default=3100.0)
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--read-subs", type=int, default=4)
    ap.add_argument("--max-subs", type=int, default=3)
    ap.add_argument("--edit-delta", type=int, default=1)
    ap.add_argument("--prefix-k", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--set", action="append", default=[], help="space separated NAME=VALUE pairs of one setting")
    ap.add_argument("--pe", action="store_true")
    args = ap.parse_args()
    import torch
    from biokanga_b200 import abi
    from biokanga_b200 import lib as bkx
    from biokanga_b200 import workload as wl
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lens = wl.chrom_layout(int(args.genome_mbp * 1e6))
    d_seq, ents = wl.make_genome(lens, seed=args.seed, device=dev)
    n = int(d_seq.numel())
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    bkx.build_suffix_array_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 0)
    torch.cuda.synchronize()
    idx = bkx.Index.from_device(d_seq.data_ptr(), n, d_sa.data_ptr(), 4, ents, name="ab", device=0, prefix_k=args.prefix_k)
    if args.pe:
        d_bases, d_offs = wl.sim_pairs(d_seq, ents, args.reads // 2, args.read_len, seed=args.seed + 100,
                                       subs=tuple(range(0, args.read_subs + 1)), device=dev)
    else:
        d_bases, d_offs = wl.sim_reads(d_seq, ents, args.reads, args.read_len, seed=args.seed + 100,
                                       subs=tuple(range(0, args.read_subs + 1)), device=dev)
    del d_seq, d_sa
    torch.cuda.empty_cache()
    nreads = args.reads
    d_out = torch.empty(nreads * 32, dtype=torch.uint8, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    d_pk2, d_flags = wl.pack2_device(d_bases, d_offs)
    first = None
    rows = []
    for setting in (args.set or [""]):
        pairs = [kv.split("=", 1) for kv in setting.split()]
        use_p2 = False
        for k_, v_ in pairs:
            if k_ == "INPUT":      # INPUT=packed2: the reads are also resident 2-bit packed (bkx_align_reads_device_packed2)
                use_p2 = v_ == "packed2"
            os.environ[k_] = v_
        p = idx.default_params(0, max_subs=args.max_subs, min_edit_dist=args.edit_delta)
        ms = []
        for rep in range(args.reps + 1):
            d_out.zero_()
            if use_p2:
                idx.align_device_packed2(p, d_bases.data_ptr(), d_pk2.data_ptr(), d_flags.data_ptr(), d_offs.data_ptr(), nreads,
                                         args.read_len, d_out.data_ptr(), None, ts.cuda_stream)
            else:
                idx.align_device(p, d_bases.data_ptr(), d_offs.data_ptr(), nreads, args.read_len, d_out.data_ptr(), None, ts.cuda_stream)
            torch.cuda.synchronize()
            if rep:
                ms.append(idx.last_kernel_ms())
        res = d_out.cpu().numpy().view(abi.RESULT_DTYPE).copy()
        if first is None:
            first = res
        same = bool(res.tobytes() == first.tobytes())
        if not same:   # which records differ, and how
            bad = np.nonzero(res != first)[0]
            print("  %d of %d records differ; by field: %s" % (
                len(bad), nreads, {f: int((res[f] != first[f]).sum()) for f in res.dtype.names if (res[f] != first[f]).any()}), flush=True)
            for i in bad[:12]:
                print("  read %d\n    first: %s\n    this : %s" % (i, first[i], res[i]), flush=True)
        row = {"setting": setting, "ms_best": min(ms), "ms_median": float(np.median(ms)), "reads_per_s": nreads / (min(ms) / 1e3),
               "identical_to_first": same}
        rows.append(row)
        print(json.dumps(row), flush=True)
        for k_, _ in pairs:
            os.environ.pop(k_, None)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ab_kernel.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
