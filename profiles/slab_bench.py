#!/usr/bin/env python
"""Times BASELINE config M (Middlebury-shaped 1984x2880, D=640, ONE pair) slab-sharded over the
ranks of a torchrun launch: every rank owns D/G disparities, two NCCL all-reduces (min, sum)
merge the AML statistics, one int64 all-reduce(min) merges the census WTA.  CUDA events, max
over ranks; prints one JSON line on rank 0.  Not the headline metric (bench.py is) -- evidence
for DESIGN.md section 7.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 profiles/slab_bench.py [--steps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--H", type=int, default=1984)
    ap.add_argument("--W", type=int, default=2880)
    ap.add_argument("--D", type=int, default=640)
    ap.add_argument("--row-bands", type=int, default=0, help="exchange mode: row bands (0 = default_row_bands: slabs stay "
                    ">= 160 disparities wide, the other ranks split the rows; 1 = disparity slabs only)")
    ap.add_argument("--mode", default="exchange", choices=["exchange", "phases"],
                    help="exchange: msn_ms_slab_fused_dev (minima / denominators traded inside the kernel over "
                         "peer memory); phases: phase A/B/C kernels around two NCCL all-reduces")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    __graft_entry__.build()
    from msnets_b200 import sharding
    from tests._synth import bordered_pair
    B = 10
    L, R = bordered_pair(args.H, args.W, 99, border=B, shift=13)
    l, r = torch.from_numpy(L[None]).cuda(), torch.from_numpy(R[None]).cuda()
    if args.mode == "exchange":
        bands = args.row_bands or sharding.ExchangeSlabMSFeatures.default_row_bands(args.D, world)
        ex = sharding.ExchangeSlabMSFeatures(1, args.H + 2 * B, args.W + 2 * B, maxdisp=args.D, row_bands=bands, board_h=B,
                                             board_w_left=B, board_w_right=B)
        grp = ex.slab_group
    else:
        bands, grp = 1, None
        ex = sharding.SlabShardedMSFeatures(1, args.H + 2 * B, args.W + 2 * B, maxdisp=args.D, board_h=B, board_w_left=B,
                                            board_w_right=B)
    out = torch.empty(ex.shape, dtype=torch.float32, device="cuda")

    parts = ex.empty_wta_parts() if args.mode == "exchange" else None

    def step():
        if args.mode == "exchange":      # WTA / second-min triples: by-product of the slab kernel, merged over the band's ranks
            ex(l, r, out=out, wta=parts)
            am, m1, m2 = sharding.slab_wta_merge(*parts, group=grp)
            return am[0, 0], m1[0, 0]
        ex(l, r, out=out)
        return sharding.slab_wta(out[0, 0], ex.d_begin, layout="dhw")   # census channel: argmin over all ranks

    step()
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        amin, vmin = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # sanity: every AML column sums to 1 over ALL ranks' disparities where the pixel has a valid cost
    s = out[0, 4:8].sum(1)
    if world > 1 and (args.mode != "exchange" or ex.slabs > 1):
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=grp)
    valid = s > 0.5          # (pixels without any valid cost have all-zero AML columns)
    err = float((s[valid] - 1).abs().max())
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if rank == 0:
        vox = args.D * args.H * args.W
        print(json.dumps({"workload": "config M: %dx%d D=%d, 1 pair, disparity-slab sharded x%d" % (
            args.H, args.W, args.D, world), "n_gpus": world, "ms_per_pair": round(float(t.item()), 2),
            "pairs_per_s": round(1e3 / float(t.item()), 3),
            "output_GB_total": round(32.0 * vox / 1e9, 1), "output_GBps_aggregate": round(32.0 * vox / float(t.item()) / 1e6, 1),
            "mode": args.mode, "sub_slabs_per_rank": getattr(ex, "subs", None), "row_bands": bands,
            "slabs": getattr(ex, "slabs", world), "per_rank_volume": list(ex.shape),
            "per_gpu_frac_of_hbm_peak": round(32.0 * vox / world / float(t.item()) / 1e6 / peak, 3),
            "collectives_per_pair": ("in-kernel exchange of 2 x [4,h,w] f32 over peer memory; " if args.mode == "exchange"
                                     else "all_reduce(min) + all_reduce(sum) over [4,h,w] f32; ") +
                                    "all_reduce(min) over [h,w] i64 keys (census WTA)",
            "aml_column_sum_max_err": err, "wta_shape": list(amin.shape)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
