#!/usr/bin/env python
"""BASELINE config #5 at size -- the federated-model proxy -- end to end on N GPUs (developer/evidence tool):

    float updates [P][25M] -> fixed point (2^-24) -> ChaCha mask -> packed Shamir k=3/n=7 (t=4) shares summed per clerk
    without being materialised (fused kernel) -> (N>1: one NCCL reduce of the clerk sums and of the mask sums)
    -> reveal from the 7 clerks -> re-expand and sum the participants' mask seeds -> unmask -> mean as floats.

Participants are sharded over the ranks (`--participants` each), walked in resident tiles of `--tile`; the synthetic
updates are drawn on the device per tile (untimed) and their exact column sums are kept in float64 to check the result.
Timed with CUDA events per stage on the context's stream, max over ranks.  One JSON line on stdout.

    python tools/federated_bench.py --participants 256
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/federated_bench.py --participants 1024
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sda_b200  # noqa: E402
from sda_b200 import LinearMaskingScheme as LMS  # noqa: E402
from sda_b200 import multi, params  # noqa: E402

FRAC = 24


def seed(tag):
    return hashlib.sha256(tag.encode()).digest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--participants", type=int, default=256, help="per GPU")
    ap.add_argument("--tile", type=int, default=64, help="resident participants per tile")
    ap.add_argument("--dim", type=int, default=25_000_000)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")      # NCCL prints its banner to fd 1
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.zeros(1, dtype=torch.int64, device="cuda")
        dist.all_reduce(warm)                  # communicator set-up (about a second) stays out of the timed reduce
        torch.cuda.synchronize()
    ctx = sda_b200.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    p = params.P61
    scheme = params.config5()
    n, dim, P, Pt = scheme.output_size(), args.dim, args.participants, args.tile
    B = scheme.batches(dim)
    ms = LMS.ChaCha(p, dim, 128)
    words = 4
    stages = {k: 0.0 for k in ("encode", "mask", "share_gen_clerk_sum", "reduce", "reveal", "mask_expand", "unmask_decode")}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        ctx.synchronize()
        stages[name] += a.elapsed_time(b)

    with torch.cuda.stream(stream):
        d_x = torch.empty((Pt, dim), dtype=torch.float32, device="cuda")
        d_q = torch.empty((Pt, dim), dtype=torch.int64, device="cuda")
        d_m = torch.empty((Pt, dim), dtype=torch.int64, device="cuda")
        d_sum = torch.zeros((n, B), dtype=torch.int64, device="cuda")
        d_seedw = torch.zeros((P, words), dtype=torch.int64, device="cuda")
        truth = torch.zeros(dim, dtype=torch.float64, device="cuda")
        gen = torch.Generator(device="cuda")
        for t0 in range(0, P, Pt):
            pt = min(Pt, P - t0)
            gen.manual_seed(1234 + rank * 100003 + t0)
            d_x[:pt].normal_(generator=gen)
            truth += d_x[:pt].sum(dim=0, dtype=torch.float64)
            stream.synchronize()
            timed("encode", lambda: ctx.fixed_encode_dev(p, FRAC, d_x, pt * dim, d_q))

            def mask_all():
                for i in range(pt):
                    ctx.mask_dev(ms, d_q[i], dim, seed(f"fed/mask/{rank}/{t0 + i}"), d_seedw[t0 + i], d_m[i])
            timed("mask", mask_all)
            seeds = b"".join(seed(f"fed/share/{rank}/{t0 + i}") for i in range(pt))
            timed("share_gen_clerk_sum",
                  lambda: ctx.share_generate_combine_dev(scheme, d_m, dim, pt, dim, seeds, d_sum, d_acc_in=d_sum if t0 else None))
        del d_x, d_q, d_m

        # the participants' mask seeds are re-expanded where they were drawn; both sums cross the ranks once
        d_mask = torch.empty(dim, dtype=torch.int64, device="cuda")
        timed("mask_expand", lambda: ctx.mask_combine_dev(ms, d_seedw, P, words, d_mask))
        d_tot = torch.empty((n, B), dtype=torch.int64, device="cuda")
        d_mtot = torch.empty(dim, dtype=torch.int64, device="cuda")

        def reduce_all():
            r1 = multi.reduce_partial_sums(d_sum, p, dst=0, final_mod=lambda t: ctx.mod_reduce_dev(p, t, n * B, d_tot, unsigned=True))
            r2 = multi.reduce_partial_sums(d_mask, p, dst=0, final_mod=lambda t: ctx.mod_reduce_dev(p, t, dim, d_mtot, unsigned=True))
            return r1, r2
        timed("reduce", reduce_all)
        if world > 1:
            dist.reduce(truth, dst=0)

        ok, err = None, None
        if rank == 0:
            d_rec = torch.empty(dim, dtype=torch.int64, device="cuda")
            timed("reveal", lambda: ctx.secret_reconstruct_dev(scheme, dim, list(range(n)), d_tot, B, n, B, d_rec))
            d_mean = torch.empty(dim, dtype=torch.float32, device="cuda")

            def finish():
                ctx.unmask_dev(ms, d_mtot, d_rec, dim, d_rec)
                ctx.fixed_decode_dev(p, FRAC, world * P, d_rec, dim, d_mean)
            timed("unmask_decode", finish)
            err = float((d_mean.double() - truth / (world * P)).abs().max())
            ok = err < 2.0 ** -(FRAC - 1)          # one rounding step per update, averaged, + float32 output rounding

    t = torch.tensor([sum(stages.values())] + [stages[k] for k in stages], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        total_ms = float(t[0])
        line = {"workload": "config#5 federated proxy: f32 -> fixed point -> ChaCha mask -> packed Shamir k=3/n=7 t=4 (fused clerk sums) "
                            "-> reveal -> unmask -> mean", "n_gpus": world, "participants_total": world * P, "dim": dim,
                "elements_per_s": world * P * dim / (total_ms * 1e-3), "ms_total": total_ms,
                "ms_per_stage_max_over_ranks": {k: float(v) for k, v in zip(stages, t[1:])},
                "max_abs_error_of_mean": err, "correct": bool(ok),
                "note": "device-resident, CUDA events per stage on the context's stream; synthetic updates drawn per tile, untimed"}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0 if (ok is None or ok) else 1


if __name__ == "__main__":
    sys.exit(main())
