#!/usr/bin/env python
"""Per-kernel measurements of every hot-path row at BASELINE-like sizes on one
B200: device-resident inputs, CUDA events on the context's stream, best and median of `--reps`
launches after 3 warm-ups, algorithmic bytes / time against the measured HBM copy peak.
Developer/evidence tool: `python tools/kernel_bench.py > gpurun_out/kernels.jsonl` (one JSON line each).
Working sets are several GB per launch (>> 126 MB L2), so there is no cache flush between launches; the few
rows whose working set fits the L2 (codec, additive reconstruct) say so in `working_set_vs_l2`."""
import argparse
import hashlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import sda_b200  # noqa: E402
from sda_b200 import LinearMaskingScheme as LMS  # noqa: E402
from sda_b200 import LinearSecretSharingScheme as LSS  # noqa: E402
from sda_b200 import params  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = sda_b200.Context(0, rng_rounds=args.rounds)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peak = peak_gbs()
    P61 = params.P61

    def seeds(tag, n):
        return b"".join(hashlib.sha256(b"%s/%d" % (tag.encode(), i)).digest() for i in range(n))

    def timeit(name, fn, elements, alg_bytes, note=""):
        if args.only and args.only not in name:
            return
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
            ctx.synchronize()
            # calls that take well under a millisecond are queued `inner` at a time between the two events, so the
            # host's preparation of one call runs under the kernel of the previous one instead of being timed
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            ctx.synchronize()
            inner = 20 if a.elapsed_time(b) < 0.5 else 1
            ts = []
            for _ in range(args.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(inner):
                    fn()
                b.record(stream)
                ctx.synchronize()
                ts.append(a.elapsed_time(b) / inner)
        best, med = min(ts), statistics.median(ts)
        line = {"kernel": name, "ms_best": best, "ms_median": med, "elements": elements,
                "elements_per_s": elements / (med * 1e-3), "algorithmic_bytes": alg_bytes,
                "GBps": alg_bytes / (med * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (med * 1e-3) / 1e9 / peak,
                "peak_GBps": peak, "variant": ctx.last_kernel() if name.startswith(("additive_split", "packed_share")) else "",
                "rng_rounds": args.rounds, "calls_per_timing": inner,
                "working_set_vs_l2": "larger than the 126 MB L2" if alg_bytes > 126e6 else "FITS the L2: not an HBM measurement",
                "note": note}
        print(json.dumps(line), flush=True)

    def empty(*shape):
        return torch.empty(shape, dtype=torch.int64, device="cuda")

    with torch.cuda.stream(stream):
        # ---- config #2: additive n=3, dim=1M, 1024 participants ------------------------------------
        s2, P, dim = params.config2(), 1024, 1_000_000
        sec = empty(P, dim)
        ctx.synth_fill_dev(2, P61, 0, P * dim, sec)
        sh = empty(P, 3, dim)
        sd = seeds("c2", P)
        timeit("additive_split cfg2 [1024][1M] n=3", lambda: ctx.share_generate_dev(s2, sec, dim, P, dim, sd, sh),
               P * dim, P * dim * 8 * 4, "read 8 + write 24 B per secret")
        out = empty(dim)
        timeit("combine cfg2 clerk job [1024][1M] (strided view of the shares)",
               lambda: ctx.share_combine_dev(s2, sh[:, 0, :], 3 * dim, P, dim, out), P * dim, P * dim * 8 + dim * 8)
        timeit("additive_reconstruct 3 x [1M]", lambda: ctx.share_combine_dev(s2, sh[0], dim, 3, dim, out), dim,
               4 * dim * 8, "launch-bound at this size")
        del sec, sh, out
        torch.cuda.empty_cache()

        # ---- configs #4 / #5: packed share generation -------------------------------------------------
        for name, mk, P in (("cfg3 k=3 n=5 t=2", params.config3, 128), ("cfg4 k=5 n=9 t=4", params.config4, 128),
                            ("cfg5 k=3 n=7 t=4", params.config5, 128)):
            s, dim = mk(), 10_000_000
            n, B = s.output_size(), s.batches(dim)
            sec = empty(P, dim)
            ctx.synth_fill_dev(4, P61, 0, P * dim, sec)
            sh = empty(P, n, B)
            sd = seeds(name, P)
            for path, label in ((2, "tensor cores"), (4, "tensor cores, share count at run time"), (1, "CUDA cores")):
                ctx.set_packed_path(path)
                timeit(f"packed_share {name} [{P}][10M] ({label})",
                       lambda: ctx.share_generate_dev(s, sec, dim, P, dim, sd, sh), P * dim, P * (dim + n * B) * 8,
                       f"8(1+n/k) = {8 * (1 + n / s.input_size()):.2f} B per secret")
            ctx.set_packed_path(0)
            # share generation fused with the clerk sums (TMEM-accumulated over the participants): both generations
            accs = empty(n, B)
            for path, label in ((0, "paired tiles"), (3, "first generation")):
                ctx.set_packed_path(path)
                timeit(f"share_generate_combine {name} [{P}][10M] ({label})",
                       lambda: ctx.share_generate_combine_dev(s, sec, dim, P, dim, sd, accs), P * dim, (P * dim + n * B) * 8,
                       "reads 8 B per secret, writes the n clerk sums once")
            ctx.set_packed_path(0)
            del accs
            if "cfg3" in name or "cfg5" in name:
                # the participant's two steps (participate.rs:53-54, :75-76): Full mask then share generation, as two
                # entry points (P mask calls, masked secrets through HBM) and as the one fused kernel
                ms = LMS.Full(P61)
                masks, masked = empty(P, dim), empty(P, dim)
                msd = seeds(name + "/mask", P)
                alg = P * (2 * dim + n * B) * 8            # secrets in, masks and shares out

                def two_steps():
                    for pi in range(P):
                        ctx.mask_dev(ms, sec[pi], dim, msd[32 * pi:32 * pi + 32], masks[pi], masked[pi])
                    ctx.share_generate_dev(s, masked, dim, P, dim, sd, sh)
                timeit(f"full mask, then packed_share {name} [{P}][10M] (P + 1 calls, masked secrets through HBM)", two_steps,
                       P * dim, alg)
                timeit(f"mask+packed_share {name} [{P}][10M] (one kernel, masked secrets in shared memory only)",
                       lambda: ctx.mask_share_generate_dev(ms, s, sec, dim, P, dim, msd, sd, masks, sh), P * dim, alg)
                del masks, masked
            if "cfg4" in name:
                out = empty(B)
                timeit("combine cfg4 clerk job [128][2M] (strided view)",
                       lambda: ctx.share_combine_dev(s, sh[:, 0, :], n * B, P, B, out), P * B, P * B * 8 + B * 8)
                rec = empty(dim)
                timeit("packed_reconstruct cfg4 9 x [2M] -> [10M]",
                       lambda: ctx.secret_reconstruct_dev(s, dim, list(range(n)), sh[0], B, n, B, rec), dim,
                       (n * B + dim) * 8, "read 8 m' + write 8 k B per batch")
                del out, rec
            del sec, sh
            torch.cuda.empty_cache()

        # ---- the same shape over a 61-bit prime that is not of Mersenne form (generic reduction path) ----------
        try:
            from oracle import oracle as O          # only to find roots of unity for the other prime
            pg = params.P61_GENERIC
            qs = [q for q in range(7, 200) if (pg - 1) % q == 0 and all(q % d for d in range(2, q))]
            sg = params.LinearSecretSharingScheme.PackedShamir(3, 5, 2, pg, O.find_root_of_order(pg, qs[0]),
                                                               O.find_root_of_order(pg, qs[1]))
            P, dim = 64, 10_000_000
            n, B = 5, sg.batches(dim)
            sec = empty(P, dim)
            ctx.synth_fill_dev(9, pg, 0, P * dim, sec)
            sh = empty(P, n, B)
            sd = seeds("gen", P)
            timeit(f"packed_share k=3 n=5 t=2 [{P}][10M], non-Mersenne 61-bit prime",
                   lambda: ctx.share_generate_dev(sg, sec, dim, P, dim, sd, sh), P * dim, P * (dim + n * B) * 8)
            del sec, sh
            torch.cuda.empty_cache()
        except Exception as e:                       # no suitable orders: skip the line
            print(json.dumps({"kernel": "packed_share generic prime", "skipped": str(e)}), flush=True)

        # ---- shapes without a fully templated kernel: the paired-tile kernel with the share count at run time
        #      (packed_tc2n.cu), and with 12 rounds the run-time-shaped kernel (packed_tcg.cu) ---------------------------
        for k_, t_, n_, rounds in ((3, 2, 6, 20), (3, 2, 4, 20), (5, 4, 8, 20), (4, 2, 8, 20), (3, 3, 7, 20), (5, 4, 10, 20),
                                   (8, 8, 20, 20), (2, 5, 9, 20), (8, 1, 12, 20), (12, 3, 20, 20), (2, 10, 13, 20),
                                   (3, 2, 6, 12), (8, 8, 20, 12)):
            if args.rounds != 20:              # a sweep at another round count: the 20-round lines only, at that count
                if rounds != 20:
                    continue
                rounds = args.rounds
            s = params.LinearSecretSharingScheme.PackedShamir(k_, n_, t_, P61, params.ROOT_ORDER_31, params.ROOT_ORDER_41)
            P, dim = 64, 10_000_000
            ctx.set_rng_rounds(rounds)
            B = s.batches(dim)
            sec = empty(P, dim)
            ctx.synth_fill_dev(10, P61, 0, P * dim, sec)
            sh = empty(P, n_, B)
            sd = seeds(f"rt{k_}{t_}{n_}", P)
            timeit(f"packed_share k={k_} t={t_} n={n_} [{P}][10M] (no fully templated kernel{'' if rounds == args.rounds else f', ChaCha{rounds}'})",
                   lambda: ctx.share_generate_dev(s, sec, dim, P, dim, sd, sh), P * dim, P * (dim + n_ * B) * 8,
                   f"8(1+n/k) = {8 * (1 + n_ / k_):.2f} B per secret")
            ctx.set_rng_rounds(args.rounds)
            del sec, sh
            torch.cuda.empty_cache()
        # ---- small calls: what the per-call flag read-back costs, and what deferred checks give back ------------------------
        s3, P, dim = params.config3(), 4, 100_000
        sec = empty(P, dim)
        ctx.synth_fill_dev(12, P61, 0, P * dim, sec)
        sh = empty(P, s3.output_size(), s3.batches(dim))
        masks, masked = empty(dim), empty(dim)
        sd = seeds("small", P)
        fm = LMS.Full(P61)
        for deferred in (False, True):
            ctx.set_deferred_checks(deferred)
            tag = "deferred checks" if deferred else "flag read back per call"
            timeit(f"small call: packed_share cfg3 [4][100k] ({tag})", lambda: ctx.share_generate_dev(s3, sec, dim, P, dim, sd, sh),
                   P * dim, P * (dim + s3.output_size() * s3.batches(dim)) * 8)
            timeit(f"small call: full_mask [100k] ({tag})", lambda: ctx.mask_dev(fm, sec[0], dim, sd[:32], masks, masked), dim, dim * 24)
        ctx.set_deferred_checks(False)
        del sec, sh, masks, masked
        # ---- additive split with a share count beyond the unrolled kernels (n = 9) ---------------------------------------
        s9, P, dim = LSS.Additive(9, P61), 128, 1_000_000
        sec = empty(P, dim)
        ctx.synth_fill_dev(11, P61, 0, P * dim, sec)
        sh = empty(P, 9, dim)
        sd = seeds("add9", P)
        timeit("additive_split n=9 [128][1M] (run-time share count)", lambda: ctx.share_generate_dev(s9, sec, dim, P, dim, sd, sh),
               P * dim, P * dim * 8 * 10, "keystream-bound: 8 draws per element")
        del sec, sh
        torch.cuda.empty_cache()

        # ---- reveal with two missing clerks (k=3, t=4, n=9 over the 61-bit prime) ---------------------------------------
        s = params.LinearSecretSharingScheme.PackedShamir(3, 9, 4, P61, params.ROOT_ORDER_11, params.ROOT_ORDER_13)
        dim = 10_000_000
        B = s.batches(dim)
        rows = empty(7, B)
        ctx.synth_fill_dev(7, P61, 0, 7 * B, rows)
        rec = empty(dim)
        timeit("packed_reconstruct k=3 t=4 n=9, clerks {0..5,7} x [3.33M] -> [10M]",
               lambda: ctx.secret_reconstruct_dev(s, dim, [0, 1, 2, 3, 4, 5, 7], rows, B, 7, B, rec), dim, (7 * B + dim) * 8)
        del rows, rec
        torch.cuda.empty_cache()

        # ---- share wire codec: one clerk's share vector of config #3 (3.33M x 9 bytes) ------------------------
        nsh = 3_333_334
        shv = empty(nsh)
        ctx.synth_fill_dev(8, P61, 0, nsh, shv)
        enc = torch.empty(10 * nsh + 64, dtype=torch.uint8, device="cuda")
        ln = ctx.varint_encode_dev(shv, nsh, enc)
        dec = empty(nsh)
        timeit("varint_encode [3.33M] 61-bit shares", lambda: ctx.varint_encode_dev(shv, nsh, enc), nsh, nsh * 8 + ln,
               "read 8 B + write 9 B per share; includes the length read-back")
        timeit("varint_decode [3.33M] 61-bit shares", lambda: ctx.varint_decode_dev(enc, ln, dec, nsh), nsh, nsh * 8 + ln)
        del shv, enc, dec
        torch.cuda.empty_cache()

        # ---- masks, dim = 25M (config #5's vector) --------------------------------------------------------
        dim = 25_000_000
        sec = empty(dim)
        ctx.synth_fill_dev(5, P61, 0, dim, sec)
        mask, masked, back = empty(dim), empty(dim), empty(dim)
        full, cc = LMS.Full(P61), LMS.ChaCha(P61, dim, 128)
        seed = hashlib.sha256(b"mask").digest()
        timeit("full_mask [25M]", lambda: ctx.mask_dev(full, sec, dim, seed, mask, masked), dim, dim * 8 * 3,
               "read 8 + write mask 8 + masked 8")
        timeit("chacha_mask [25M] (20 rounds, wire format)", lambda: ctx.mask_dev(cc, sec, dim, seed, mask, masked), dim,
               dim * 8 * 2, "read 8 + write masked 8; the mask is the 4 seed words")
        timeit("unmask [25M]", lambda: ctx.unmask_dev(full, mask, masked, dim, back), dim, dim * 8 * 3)
        Pm = 256
        seeds_t = torch.randint(0, 1 << 32, (Pm, 4), dtype=torch.int64, device="cuda")
        timeit(f"chacha_mask_combine {Pm} seeds x [25M]", lambda: ctx.mask_combine_dev(cc, seeds_t, Pm, 4, back), Pm * dim,
               dim * 8, "compute-bound by construction: P keystream blocks per 8 outputs, elements = draws")
        masks = empty(64, dim)
        ctx.synth_fill_dev(6, P61, 0, 64 * dim, masks)
        timeit("full_mask_combine [64][25M]", lambda: ctx.mask_combine_dev(full, masks, 64, dim, back), 64 * dim,
               65 * dim * 8)


if __name__ == "__main__":
    main()
