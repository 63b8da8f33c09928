#!/usr/bin/env python
"""bench.py -- field-elements/s of the SDA hot path (share-gen + clerk-sum) on N B200s.

Workload (BASELINE.json config #3, the configuration the share-gen target is quoted on):
packed Shamir k=3 / n=5 (t=2) over p = 2^61-1, dim = 10M secrets per participant.  One *step*
is one pass of the hot path over one resident tile of T participants per GPU:
    K2  sda_share_generate_dev   secrets[T][dim]      -> shares[T][5][B]      (B = ceil(dim/3))
    K3  sda_share_combine_dev x5 shares[T][c][B]      -> clerk_sum[c][B]      (one per clerk)
    N>1 sda_partial_sums_reduce_dev: one ncclReduce (uint64 sum) of clerk_sum[5][B] over the ranks + one mod-p pass
        on rank 0, inside the library (torch.distributed only launches the ranks, hands rank 0's NCCL id to the
        others and takes the max of the timings)
(participants are sharded across GPUs: weak scaling, no other data-path collective).
`kernels.config4_clerk_sum` / `kernels.config5_e2e` are BASELINE configs #4 and #5 at their named sizes per GPU.
`value` = secrets processed by all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`   = the same metric through the host-buffer C-ABI calls a Rust shim would make
          (sda_share_generate per participant, sda_share_combine per clerk), pinned host buffers,
          H2D and D2H inside the timed region.
`--impl reference` times the reference's own CPU algorithm (the oracle's literal restatement:
per-batch Newton interpolation as tss 0.2 does, signed i64 `%`) on all host cores.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "field-elements/sec (share-gen + clerk-sum)"
UNIT = "field-elements/s"
DIM = 10_000_000
K, T_PRIV, N_SHARES = 3, 2, 5


def seeds_for(step, rank, count):
    return b"".join(hashlib.sha256(b"bench/%d/%d/%d" % (step, rank, i)).digest() for i in range(count))


def workload_config(args, world):
    return {
        "workload": "config#3 packed Shamir k=3/n=5 t=2, p=2^61-1, dim=10M: share-gen + per-clerk combine"
                    + (" + NCCL reduce of clerk sums" if world > 1 else ""),
        "dim": DIM, "participants_per_gpu_per_step": args.participants, "participants_total_config": 4096,
        "prime_modulus": (1 << 61) - 1, "omega_secrets": "order 7", "omega_shares": "order 11",
        "rng": f"ChaCha{args.rounds} keystream, rand-0.3 gen_range", "parallelism": f"participants sharded x{world}",
        "share_gen_kernel": {"auto": "tcgen05 byte-limb GEMM, paired tiles", "tc": "tcgen05 byte-limb GEMM, paired tiles",
                             "tc1": "tcgen05 byte-limb GEMM, first generation", "cuda": "IMAD.WIDE CUDA cores"}[args.packed_path],
        "l2": "inputs (>= 10 GB per step) far exceed the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's literal restatement on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_step(O, scheme_o, modulus, dim, participants, threads, tag, os_rng=False):
    """generate + per-clerk combine for `participants` vectors of `dim`, one participant per thread.
    os_rng: draw the sharing randomness the way the reference does (rand 0.3 `OsRng`: one getrandom(2) per word,
    additive.rs:17,43 / packed_shamir.rs:40-43) instead of from the injected ChaCha20 stream of the parity mode."""
    import numpy as np
    n = scheme_o.share_count
    B = (dim + scheme_o.secret_count - 1) // scheme_o.secret_count
    secrets = [O.synth_fill(3, modulus, i * dim, dim) for i in range(participants)]
    shares = [None] * participants
    t0 = time.perf_counter()

    def work(i):
        rng = O.rng_os() if os_rng else O.rng_from_seed_bytes(hashlib.sha256(b"%s/%d" % (tag.encode(), i)).digest())
        shares[i] = O.share_generate(scheme_o, secrets[i], rng)

    ths = []
    for base in range(0, participants, threads):
        ths = [threading.Thread(target=work, args=(i,)) for i in range(base, min(participants, base + threads))]
        [t.start() for t in ths]
        [t.join() for t in ths]
    stacked = np.stack(shares)          # [P][n][B]

    def comb(c):
        O.share_combine(modulus, np.ascontiguousarray(stacked[:, c, :]))

    ths = [threading.Thread(target=comb, args=(c,)) for c in range(n)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    assert stacked.shape == (participants, n, B)
    return dt


def oracle_scheme(O):
    from sda_b200 import params
    c = params.config3().c
    return O.SharingScheme(c.kind, c.share_count, c.secret_count, c.privacy_threshold, c.modulus, c.omega_secrets,
                           c.omega_shares), c.modulus


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.build()
    so, m = oracle_scheme(O)
    cores = os.cpu_count() or 1
    dim_s = args.ref_dim
    # calibrate so that one step is about ref_seconds
    t = cpu_step(O, so, m, min(dim_s, 50_000), cores, cores, "cal")
    per_el = t / (min(dim_s, 50_000) * cores)
    dim_s = int(max(30_000, min(DIM, args.ref_seconds / (per_el * cores))))
    for w in range(args.warmup):
        cpu_step(O, so, m, dim_s, cores, cores, f"w{w}")
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_step(O, so, m, dim_s, cores, cores, f"s{s}")
    dt = time.perf_counter() - t0
    value = args.steps * cores * dim_s / dt
    sample = f"{cores} participants x {dim_s} secrets per step (of dim 10M), one participant per thread"
    # the same loop with the reference's own randomness source (one getrandom(2) per draw), one bounded step
    dim_os = max(10_000, dim_s // 4)
    t_os = cpu_step(O, so, m, dim_os, cores, cores, "os", os_rng=True)
    faithful = {"value": cores * dim_os / t_os, "unit": UNIT, "cores": cores, "kind": "port",
                "rng": "OsRng: one getrandom(2) per 32-bit word, as rand 0.3 does (additive.rs:17,43, packed_shamir.rs:40-43)",
                "sample": f"{cores} participants x {dim_os} secrets, one step"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i64 (mod 2^61-1, 128-bit products)", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "rng": "injected ChaCha20 stream (the parity mode)", "faithful_os_rng": faithful},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = snipsco/sda's CPU algorithm restated in C (oracle/sda_oracle.c: per-batch Newton "
                "interpolation like tss 0.2, widened to 128-bit products for the 61-bit prime; injected ChaCha20 "
                "instead of one getrandom(2) per draw). The Rust crate itself cannot be built in this image.",
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and clock-event reasons of one GPU, polled every 5 ms from a thread of this process through
    NVML (pynvml) from before the warm-up; stop(t0, t1) keeps the samples whose timestamps fall inside the timed
    region [t0, t1] (wall clock).  If NVML cannot be loaded the same query runs as an `nvidia-smi -lms 25` child.
    A region too short to hold two samples falls back to every sample taken under load (warm-up + timed steps run
    the same kernels back to back) and says so in `window`."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, pci_bus_id=None):
        self.rows = []                 # (wall time, sm MHz, max sm MHz, power W, [reasons])
        self.p = self.f = self.thread = None
        self.source = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if pci_bus_id:
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap),
                     ("hw_power_brake_slowdown", pynvml.nvmlClocksEventReasonHwPowerBrakeSlowdown))

            def poll():
                while not self._stop.is_set():
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.rows.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), sm_max,
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, [n for n, bit in names if mask & bit]))
                    except Exception:
                        pass
                    self._stop.wait(0.005)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "NVML, 5 ms polling"
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi -lms 25"
        except OSError:
            self.p = None

    def ready(self, timeout=3.0):
        """block until the first sample exists (nvidia-smi needs a moment to start; NVML is immediate)"""
        t_end = time.time() + timeout
        while time.time() < t_end:
            if self.rows or (self.f is not None and os.path.getsize(self.f.name) > 0):
                return True
            time.sleep(0.01)
        return False

    def _collect_smi(self):
        import datetime
        if self.p is None:
            return
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]),
                                  [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                                     parts[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        else:
            self._collect_smi()
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 2:          # use everything under load since the warm-up
            inside = [r for r in rows if r[0] <= t1 and r[3] > 0.5 * max(x[3] for x in rows)] if rows else []
            window = "warm-up + timed region (samples drawing > half of the maximum power)"
        if inside:
            out.update(sm_mhz=statistics.median(r[1] for r in inside), sm_max_mhz=max(r[2] for r in inside),
                       reasons=sorted({n for r in inside for n in r[4]}), samples=len(inside),
                       power_w_max=max(r[3] for r in inside), window=window)
        return out


# ---------------------------------------------------------------------------------------------------
# BASELINE configs #4 and #5 at their named sizes, through the same context (reported under `kernels`)
# ---------------------------------------------------------------------------------------------------
def run_config4(ctx, stream, torch, dist, world, rank, peak):
    """config #4: one clerk's job of the k=5/n=9 scheme, [65536][2M] share rows (1.05 TB), participants sharded over the
    ranks; every rank walks its 65536/N rows in resident tiles with sda_share_combine_dev (running sum), then ONE
    sda_partial_sums_reduce_dev.  The synthetic tile (identical on every rank, filled untimed) is re-read for every tile
    of the shard: 8 B of HBM traffic per share element either way."""
    from sda_b200 import params
    s4 = params.config4()
    p, L, rows_total = s4.modulus, 2_000_000, 65536
    rows_rank = rows_total // world
    R = min(rows_rank, 4096)                      # 65.5 GB resident
    tiles = rows_rank // R
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    with torch.cuda.stream(stream):
        d_tile = torch.empty((R, L), dtype=torch.int64, device="cuda")
        d_acc = torch.empty(L, dtype=torch.int64, device="cuda")
        d_one = torch.empty(L, dtype=torch.int64, device="cuda")
        ctx.synth_fill_dev(4, p, 0, R * L, d_tile)
        ctx.share_combine_dev(s4, d_tile, L, R, L, d_one)            # warm-up and the check's column sums of one tile
        if world > 1:
            ctx.partial_sums_reduce_dev(p, d_acc.zero_(), L, 0)      # communicator warm-up
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        a, b, c = ev(), ev(), ev()
        a.record(stream)
        for t in range(tiles):
            ctx.share_combine_dev(s4, d_tile, L, R, L, d_acc, d_acc_in=d_acc if t else None)
        b.record(stream)
        if world > 1:
            ctx.partial_sums_reduce_dev(p, d_acc, L, 0)
        c.record(stream)
        ctx.synchronize()
        ms = torch.tensor([a.elapsed_time(c), b.elapsed_time(c)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ok = None
        if rank == 0:   # sampled columns: total == (tiles * world) * (column sum of one tile) mod p, in Python integers
            cols = [0, 1, 77, L // 2, L - 1]
            one = [int(d_one[cidx]) for cidx in cols]
            got = [int(d_acc[cidx]) for cidx in cols]
            ok = got == [(v * tiles * world) % p for v in one]
        del d_tile, d_acc, d_one
    torch.cuda.empty_cache()
    total_ms, reduce_ms = float(ms[0]), float(ms[1])
    return {"ms": total_ms, "share_elements_per_s": rows_total * L / (total_ms * 1e-3), "rows_per_gpu": rows_rank,
            "rows_total": rows_total, "L": L, "resident_tile_rows": R, "GBps_per_gpu": rows_rank * L * 8 / (total_ms * 1e-3) / 1e9,
            "frac_of_hbm_per_gpu": rows_rank * L * 8 / (total_ms * 1e-3) / 1e9 / peak,
            "collective_ms": reduce_ms if world > 1 else 0.0, "collective_share_of_step": reduce_ms / total_ms if world > 1 else 0.0,
            "correct": ok, "note": "config #4 clerk sum at size: combine over the rank's rows (running sum over resident tiles) + "
                                   "one NCCL reduce of [2M] u64 + mod pass inside the library; strong scaling over N"}


def run_config5(ctx, stream, torch, dist, world, rank, participants, tile):
    """config #5, the federated proxy end to end: float updates [P][25M] -> fixed point (2^-24) -> ChaCha mask -> packed Shamir
    k=3/n=7 (t=4) shares summed per clerk in TMEM (fused kernel) -> one library-side NCCL reduce of the clerk sums and of the
    mask sums -> reveal from the 7 clerks -> unmask -> mean.  1024 participants per GPU (8192 at 8 GPUs = the named size)."""
    from sda_b200 import LinearMaskingScheme as LMS
    from sda_b200 import params
    FRAC = 24
    p = params.P61
    scheme = params.config5()
    n, dim, P, Pt = scheme.output_size(), 25_000_000, participants, tile
    B = scheme.batches(dim)
    ms_ = LMS.ChaCha(p, dim, 128)
    words = 4
    stages = {k: 0.0 for k in ("encode_mask", "share_gen_clerk_sum", "mask_expand", "reduce", "reveal", "unmask_decode")}
    sd = lambda tag: hashlib.sha256(tag.encode()).digest()   # noqa: E731

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        ctx.synchronize()
        stages[name] += a.elapsed_time(b)

    with torch.cuda.stream(stream):
        d_x = torch.empty((Pt, dim), dtype=torch.float32, device="cuda")
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

            def encode_mask():     # one pass per participant: float -> fixed point -> + ChaCha mask (participate.rs:53-54)
                ctx.fixed_encode_mask_dev(ms_, p, FRAC, d_x, dim, b"".join(sd(f"fed/mask/{rank}/{t0 + i}") for i in range(pt)),
                                          d_seedw[t0:], d_m, P=pt)
            timed("encode_mask", encode_mask)
            seeds = b"".join(sd(f"fed/share/{rank}/{t0 + i}") for i in range(pt))
            timed("share_gen_clerk_sum",
                  lambda: ctx.share_generate_combine_dev(scheme, d_m, dim, pt, dim, seeds, d_sum, d_acc_in=d_sum if t0 else None))
        del d_x, d_m
        d_mask = torch.empty(dim, dtype=torch.int64, device="cuda")
        timed("mask_expand", lambda: ctx.mask_combine_dev(ms_, d_seedw, P, words, d_mask))

        if world > 1:            # NCCL sets up its channels for a payload size on first use: keep that out of the timed exchange
            warm = torch.zeros(max(n * B, dim), dtype=torch.int64, device="cuda")
            ctx.partial_sums_reduce_dev(p, warm, n * B, 0)
            ctx.partial_sums_reduce_dev(p, warm, dim, 0)
            ctx.synchronize()
            del warm
            dist.barrier()

        def reduce_all():
            ctx.partial_sums_reduce_dev(p, d_sum, n * B, 0)
            ctx.partial_sums_reduce_dev(p, d_mask, dim, 0)
        timed("reduce", reduce_all)
        if world > 1:
            dist.reduce(truth, dst=0)
        ok, err = None, None
        if rank == 0:
            d_rec = torch.empty(dim, dtype=torch.int64, device="cuda")
            timed("reveal", lambda: ctx.secret_reconstruct_dev(scheme, dim, list(range(n)), d_sum, B, n, B, d_rec))
            d_mean = torch.empty(dim, dtype=torch.float32, device="cuda")

            def finish():
                ctx.unmask_dev(ms_, d_mask, d_rec, dim, d_rec)
                ctx.fixed_decode_dev(p, FRAC, world * P, d_rec, dim, d_mean)
            timed("unmask_decode", finish)
            err = float((d_mean.double() - truth / (world * P)).abs().max())
            ok = err < 2.0 ** -(FRAC - 1)
            del d_rec, d_mean
        del d_sum, d_seedw, d_mask, truth
    torch.cuda.empty_cache()
    t = torch.tensor([sum(stages.values())] + [stages[k] for k in stages], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    return {"ms": total_ms, "elements_per_s": world * P * dim / (total_ms * 1e-3), "participants_total": world * P,
            "participants_per_gpu": P, "dim": dim, "ms_per_stage_max_over_ranks": {k: float(v) for k, v in zip(stages, t[1:])},
            "max_abs_error_of_mean": err, "correct": ok,
            "note": "config #5 federated proxy, device-resident, CUDA events per stage (max over ranks); the named size is "
                    "8192 participants on 8 GPUs, fewer GPUs run the same 1024 participants per GPU (weak scaling)"}


def pin_to_gpu_cores(torch, local, world):
    """Keep this rank's host threads on the cores next to its GPU (the PCIe root's local_cpulist), and when several ranks
    share that list give each its own share of it: the host side of the e2e leg is memcpy and DMA submission."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = []
        for part in open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        if world > 1 and len(allowed) >= 2 * world:
            per = len(allowed) // world
            allowed = allowed[local * per:(local + 1) * per]
        os.sched_setaffinity(0, allowed)
        return f"{allowed[0]}-{allowed[-1]} ({len(allowed)} cpus)"
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """per-launch DRAM bytes of the share-gen kernel from the committed ncu capture, if any"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import sda_b200
    from sda_b200 import params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line.  NCCL printf()s its "NCCL version ..." banner to fd 1 (this image sets
    # NCCL_DEBUG=VERSION; NCCL_DEBUG_FILE does not move it), so fd 1 is pointed at stderr for the duration of the
    # run and the JSON line goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sda_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = pin_to_gpu_cores(torch, local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = sda_b200.Context(local, rng_rounds=args.rounds)
    ctx.set_packed_path({"auto": 0, "cuda": 1, "tc": 2, "tc1": 3}[args.packed_path])
    if world > 1:
        # the library's own NCCL communicator (C ABI): rank 0 creates the id, torch.distributed only carries the 128 bytes
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(sda_b200.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init_rank(bytes(idt.cpu().tolist()), world, rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    scheme = params.config3()
    p = scheme.modulus
    T, dim, n = args.participants, DIM, N_SHARES
    B = scheme.batches(dim)

    with torch.cuda.stream(stream):
        d_sec = torch.empty((T, dim), dtype=torch.int64, device="cuda")
        d_sh = torch.empty((T, n, B), dtype=torch.int64, device="cuda")
        d_sum = torch.empty((n, B), dtype=torch.int64, device="cuda")
        d_tot = torch.empty((n, B), dtype=torch.int64, device="cuda")
        # synthetic canonical secrets, generated on the device (untimed); rank-distinct
        ctx.synth_fill_dev(3, p, rank * T * dim, T * dim, d_sec)
        ctx.synchronize()

        ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
        gen_ms, comb_ms = [], []

        def step(i, timed):
            seeds = seeds_for(i, rank, T)
            a, b, c = ev(), ev(), ev()
            a.record(stream)
            ctx.share_generate_dev(scheme, d_sec, dim, T, dim, seeds, d_sh)
            b.record(stream)
            for cl in range(n):
                ctx.share_combine_dev(scheme, d_sh[:, cl, :], n * B, T, B, d_sum[cl])
            if world > 1:
                # one collective: canonical partials < p summed as 64-bit integers (<= 8 of them cannot wrap),
                # then one mod-p pass on the root -- sda_partial_sums_reduce_dev, NCCL called by the library
                ctx.partial_sums_reduce_dev(p, d_sum, n * B, 0)
            c.record(stream)
            if timed:
                return a, b, c
            return None

        sampler = None
        if rank == 0:
            bus = None
            try:
                pr = torch.cuda.get_device_properties(local)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            except Exception:
                bus = None
            sampler = ClockSampler(local, bus)
            sampler.ready()
        for w in range(args.warmup):
            step(-1 - w, False)
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches0 = ctx.launch_count()
        t_start, t_end = ev(), ev()
        torch.cuda.synchronize()
        wall0 = time.time()
        t_start.record(stream)
        marks = [step(i, True) for i in range(args.steps)]
        t_end.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        wall1 = time.time()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop(wall0, wall1) if sampler else None
        launches = ctx.launch_count() - launches0
        total_ms = t_start.elapsed_time(t_end)
        for a, b, c in marks:
            gen_ms.append(a.elapsed_time(b))
            comb_ms.append(b.elapsed_time(c))
        kernel_name = ctx.last_kernel()

        # the same share-gen launch at the other keystream round counts the context offers (not part of
        # `value`: the sharing randomness replaces the reference's OsRng, the round count is a context option)
        other_rounds = {}
        if rank == 0 and not args.no_round_sweep:
            for r in (12, 8):
                if r == args.rounds:
                    continue
                ctx.set_rng_rounds(r)
                ctx.share_generate_dev(scheme, d_sec, dim, T, dim, seeds_for(-100 - r, rank, T), d_sh)
                ts = []
                for i in range(3):
                    a, b = ev(), ev()
                    a.record(stream)
                    ctx.share_generate_dev(scheme, d_sec, dim, T, dim, seeds_for(-200 - r - i, rank, T), d_sh)
                    b.record(stream)
                    ctx.synchronize()
                    ts.append(a.elapsed_time(b))
                other_rounds[r] = statistics.median(ts)
            ctx.set_rng_rounds(args.rounds)

        # the fused participant -> clerk kernel: same clerk sums without materialising the
        # shares, the participant sum accumulated in TMEM.  Checked against the K2 + K3 result, then timed.
        fused_ms = None
        if rank == 0 and not args.no_round_sweep:
            sd = seeds_for(-300, rank, T)
            ctx.share_generate_dev(scheme, d_sec, dim, T, dim, sd, d_sh)
            for cl in range(n):
                ctx.share_combine_dev(scheme, d_sh[:, cl, :], n * B, T, B, d_sum[cl])
            ctx.share_generate_combine_dev(scheme, d_sec, dim, T, dim, sd, d_tot)
            ctx.synchronize()
            if not torch.equal(d_tot, d_sum):
                raise SystemExit("bench self-check failed: fused share-gen+clerk-sum != share-gen then combine")
            ts = []
            for i in range(3):
                a, b = ev(), ev()
                a.record(stream)
                ctx.share_generate_combine_dev(scheme, d_sec, dim, T, dim, seeds_for(-310 - i, rank, T), d_tot)
                b.record(stream)
                ctx.synchronize()
                ts.append(a.elapsed_time(b))
            fused_ms = statistics.median(ts)

        # the participant's two steps in one kernel (participate.rs:53-54 then :75-76): Full mask + share generation, the
        # masked secrets never written.  Checked against sda_mask_dev followed by sda_share_generate_dev, then timed.
        masked = None
        if rank == 0 and world == 1 and not args.no_round_sweep and args.rounds == 20:
            Tm = min(T, 64)
            fm = sda_b200.LinearMaskingScheme.Full(p)
            msd, ssd = seeds_for(-400, rank, Tm), seeds_for(-401, rank, Tm)
            d_masks = torch.empty((Tm, dim), dtype=torch.int64, device="cuda")
            ctx.mask_share_generate_dev(fm, scheme, d_sec, dim, Tm, dim, msd, ssd, d_masks, d_sh)
            fused_kernel = ctx.last_kernel()
            d_m0, d_md0 = torch.empty(dim, dtype=torch.int64, device="cuda"), torch.empty(dim, dtype=torch.int64, device="cuda")
            d_s0 = torch.empty((n, B), dtype=torch.int64, device="cuda")
            ctx.mask_dev(fm, d_sec[Tm - 1], dim, msd[32 * (Tm - 1):32 * Tm], d_m0, d_md0)
            ctx.share_generate_dev(scheme, d_md0, dim, 1, dim, ssd[32 * (Tm - 1):32 * Tm], d_s0)
            ctx.synchronize()
            if not (torch.equal(d_masks[Tm - 1], d_m0) and torch.equal(d_sh[Tm - 1], d_s0)):
                raise SystemExit("bench self-check failed: fused mask + share generation != mask then share generation")
            ts = []
            for i in range(3):
                a, b = ev(), ev()
                a.record(stream)
                ctx.mask_share_generate_dev(fm, scheme, d_sec, dim, Tm, dim, seeds_for(-410 - i, rank, Tm), ssd, d_masks, d_sh)
                b.record(stream)
                ctx.synchronize()
                ts.append(a.elapsed_time(b))
            mms = statistics.median(ts)
            masked = {"ms": mms, "participants": Tm, "elements_per_s": Tm * dim / (mms * 1e-3),
                      "GBps": Tm * (2 * dim + n * B) * 8 / (mms * 1e-3) / 1e9, "kernel": fused_kernel,
                      "note": "sda_mask_share_generate_dev: Full mask + share generation in one kernel, masked secrets never "
                              "written to HBM (secrets in, masks and shares out); not part of `value`"}
            del d_masks, d_m0, d_md0, d_s0
            torch.cuda.empty_cache()

        # BASELINE config #2 beside it (additive 3-way split, dim 1M, 1024 participants, then one clerk's sum): the
        # additive kernels of the same path, device-resident, same timing method; reported under `kernels`, not in `value`
        cfg2 = None
        if rank == 0 and world == 1 and not args.no_round_sweep:
            s2, P2, dim2 = params.config2(), 1024, 1_000_000
            d_sec2 = torch.empty((P2, dim2), dtype=torch.int64, device="cuda")
            d_sh2 = torch.empty((P2, 3, dim2), dtype=torch.int64, device="cuda")
            d_sum2 = torch.empty(dim2, dtype=torch.int64, device="cuda")
            ctx.synth_fill_dev(2, p, 0, P2 * dim2, d_sec2)
            ts_split, ts_comb = [], []
            for i in range(4):
                a, b, c = ev(), ev(), ev()
                a.record(stream)
                ctx.share_generate_dev(s2, d_sec2, dim2, P2, dim2, seeds_for(-400 - i, rank, P2), d_sh2)
                b.record(stream)
                ctx.share_combine_dev(s2, d_sh2[:, 0, :], 3 * dim2, P2, dim2, d_sum2)
                c.record(stream)
                ctx.synchronize()
                if i:                                   # first iteration is the warm-up
                    ts_split.append(a.elapsed_time(b))
                    ts_comb.append(b.elapsed_time(c))
            ms_split, ms_comb = statistics.median(ts_split), statistics.median(ts_comb)
            cfg2 = {"config2_additive_split": {"ms": ms_split, "elements_per_s": P2 * dim2 / (ms_split * 1e-3),
                                               "GBps": P2 * dim2 * 32 / (ms_split * 1e-3) / 1e9, "rng_rounds": args.rounds,
                                               "note": "[1024][1M] secrets -> 3 shares each: reads 8 B, writes 24 B per element"},
                    "config2_clerk_combine": {"ms": ms_comb, "share_elements_per_s": P2 * dim2 / (ms_comb * 1e-3),
                                              "GBps": (P2 + 1) * dim2 * 8 / (ms_comb * 1e-3) / 1e9,
                                              "note": "one clerk's [1024][1M] job (strided view of the shares)"}}
            del d_sec2, d_sh2, d_sum2

        # correctness spot check of what was just timed (cheap, outside the timed region):
        # reveal(clerk sums) == column sums of the secrets
        if rank == 0 and world == 1:
            d_rec = torch.empty(dim, dtype=torch.int64, device="cuda")
            d_ref = torch.empty(dim, dtype=torch.int64, device="cuda")
            ctx.secret_reconstruct_dev(scheme, dim, list(range(n)), d_sum, B, n, B, d_rec)
            ctx.share_combine_dev(scheme, d_sec, dim, T, dim, d_ref)
            ctx.synchronize()
            if not torch.equal(d_rec, d_ref):
                raise SystemExit("bench self-check failed: reveal(clerk sums) != sum of secrets")
            del d_rec, d_ref

    t_ms = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    value = world * T * dim * args.steps / (total_ms * 1e-3)

    # ---- e2e: host-buffer C-ABI calls, copies inside the timed region ----------------------------
    # The calls a Rust shim would make, on host vectors: sda_share_generate per participant, sda_share_combine_rows per
    # clerk.  One call is PCIe work (80 MB in, 133 MB out around a 50 us kernel); the library's concurrency model is one
    # context per client thread, so the leg runs `--e2e-threads` client threads, each with its own context (own streams and
    # device buffers): while one thread's call is still copying shares out, the next thread's call is copying secrets in,
    # and both directions of the link stay busy.  Timed for at least a second; pinned buffers (sda_host_alloc) are the
    # headline, ordinary pageable numpy arrays are reported beside it.
    Te = 0 if args.no_e2e else args.e2e_participants
    e2e_value = e2e_pageable = pcie = None
    e2e_steps = 0
    if Te > 0:
        from concurrent.futures import ThreadPoolExecutor
        nthreads = max(1, args.e2e_threads)
        workers = [sda_b200.Context(local, rng_rounds=args.rounds) for _ in range(nthreads)]
        for w in workers:
            w.set_packed_path({"auto": 0, "cuda": 1, "tc": 2, "tc1": 3}[args.packed_path])
        pool = ThreadPoolExecutor(nthreads)
        h_sec = [ctx.pinned_empty(dim) for _ in range(Te)]
        h_sh2 = [ctx.pinned_empty(Te * n * B).reshape(Te, n, B) for _ in range(2)]
        h_out = ctx.pinned_empty(n * B).reshape(n, B)
        for i in range(Te):
            h_sec[i][:] = d_sec[i].cpu().numpy()

        def run_step(i, sec, sh_new, sh_prev, out):
            """the participants of round i share (sda_share_generate, one call each) while the clerks sum the shares of
            round i - 1 (sda_share_combine_rows, one call per clerk): the participants' calls are copy-out heavy, the clerks'
            copy-in heavy, so the two kinds of client keep both directions of the link busy"""
            seeds = seeds_for(1000 + i, rank, Te)

            def gen(q):
                workers[q % nthreads].share_generate(scheme, sec[q], seeds[32 * q:32 * q + 32], out=sh_new[q])

            def comb(cl):
                # the clerk receives its column of every participation (server snapshot transpose,
                # snapshot.rs:11-27): a `Vec<Vec<Share>>` of P rows, passed as row pointers
                workers[(Te + cl) % nthreads].share_combine(scheme, [sh_prev[q, cl] for q in range(Te)], out=out[cl])
            if args.e2e_mode == "phases":      # participants first, then the clerks (of the previous round's shares)
                list(pool.map(gen, range(Te)))
                if sh_prev is not None:
                    list(pool.map(comb, range(n)))
                return
            futs = [pool.submit(gen, q) for q in range(Te)] + ([pool.submit(comb, cl) for cl in range(n)] if sh_prev is not None else [])
            for f_ in futs:
                f_.result()

        run_step(-1, h_sec, h_sh2[1], None, h_out)
        run_step(-2, h_sec, h_sh2[0], h_sh2[1], h_out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        while e2e_steps < 3 or time.perf_counter() - t0 < 1.0:
            run_step(e2e_steps, h_sec, h_sh2[(e2e_steps + 1) % 2], h_sh2[e2e_steps % 2], h_out)
            e2e_steps += 1
        e2e_s = time.perf_counter() - t0
        t_e = torch.tensor([e2e_s / e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e_value = world * Te * dim / float(t_e.item())
        # what the last timed step combined: the shares generated by the step before it
        last_prev = h_sh2[(e2e_steps - 1) % 2]
        # the same calls on pageable memory (what a plain Vec<i64> is): staged through pinned bounce buffers inside the library
        p_sec = [np.array(h) for h in h_sec]
        p_prev = np.array(last_prev)
        p_sh = np.empty((Te, n, B), dtype=np.int64)
        p_out = np.empty((n, B), dtype=np.int64)
        run_step(-3, p_sec, p_sh, p_prev, p_out)
        t0 = time.perf_counter()
        run_step(-4, p_sec, p_sh, p_prev, p_out)
        t_p = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_p, op=dist.ReduceOp.MAX)
        e2e_pageable = world * Te * dim / float(t_p.item())
        if not np.array_equal(p_out, h_out) and rank == 0:      # both summed the same shares
            raise SystemExit("bench self-check failed: pageable and pinned host paths disagree")
        # what the link itself gives this rank (pinned, 256 MB each way, alone and both ways at once)
        hb = torch.empty(32 << 20, dtype=torch.int64).pin_memory()
        hb2 = torch.empty(32 << 20, dtype=torch.int64).pin_memory()
        db, db2 = torch.empty_like(hb, device="cuda"), torch.empty_like(hb, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def link(up, down):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                if up:
                    with torch.cuda.stream(s1):
                        db.copy_(hb, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        hb2.copy_(db2, non_blocking=True)
            torch.cuda.synchronize()
            return 4 * hb.numel() * 8 / (time.perf_counter() - t0) / 1e9
        link(True, True)
        if world > 1:
            dist.barrier()
        pcie = {"h2d_GBps": link(True, False), "d2h_GBps": link(False, True), "duplex_each_way_GBps": link(True, True),
                "note": "pinned 256 MB copies on this rank, all ranks at the same time"}
        del hb, hb2, db, db2
        pool.shutdown()
        for w in workers:
            w.close()
    h2d = Te * dim * 8 + n * Te * B * 8
    d2h = Te * n * B * 8 + n * B * 8

    # ---- configs #4 and #5 at size (every rank takes part; the big buffers of the main workload are released first) ----
    peak, peak_src = load_peaks()
    cfg4 = cfg5 = None
    if not args.no_configs45:
        del d_sec, d_sh, d_sum, d_tot
        torch.cuda.empty_cache()
        cfg4 = run_config4(ctx, stream, torch, dist, world, rank, peak)
        cfg5 = run_config5(ctx, stream, torch, dist, world, rank, args.cfg5_participants, 64)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K2 share-gen) --------------------------------------------
    alg_bytes = T * (dim * 8 + n * B * 8)                       # read secrets + write shares
    gen_avg_ms = sum(gen_ms) / len(gen_ms)
    achieved = alg_bytes / (gen_avg_ms * 1e-3) / 1e9
    comb_bytes = n * (T * B * 8 + B * 8)
    comb_avg_ms = sum(comb_ms) / len(comb_ms)
    traffic = load_traffic()
    traffic_key = "packed_share_bytes_per_launch" if "tcgen05" in kernel_name else "packed_share_cuda_bytes_per_launch"
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic.get(traffic_key) if (T, args.rounds) == (256, 20) else None,
                "traffic_source": "profiles/traffic.json (ncu --set full, same launch shape)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": gen_avg_ms,
                "share_of_step": gen_avg_ms / (gen_avg_ms + comb_avg_ms)}
    kernels = {
        "share_gen": {"ms": gen_avg_ms, "elements_per_s": T * dim / (gen_avg_ms * 1e-3), "GBps": achieved,
                      "frac_of_hbm": achieved / peak, "rng_rounds": args.rounds},
        **{f"share_gen_chacha{r}": {"ms": ms, "elements_per_s": T * dim / (ms * 1e-3), "GBps": alg_bytes / (ms * 1e-3) / 1e9,
                                    "frac_of_hbm": alg_bytes / (ms * 1e-3) / 1e9 / peak, "rng_rounds": r}
           for r, ms in other_rounds.items()},
        **({"fused_share_gen_clerk_sum": {"ms": fused_ms, "elements_per_s": T * dim / (fused_ms * 1e-3),
                                          "GBps": T * dim * 8 / (fused_ms * 1e-3) / 1e9,
                                          "frac_of_hbm": T * dim * 8 / (fused_ms * 1e-3) / 1e9 / peak,
                                          "note": "sda_share_generate_combine_dev: reads 8 B per secret, shares never "
                                                  "materialised; not part of `value`"}} if fused_ms else {}),
        **({"mask_share_gen_fused": dict(masked, frac_of_hbm=masked["GBps"] / peak)} if masked else {}),
        **({k: dict(v, frac_of_hbm=v["GBps"] / peak) for k, v in cfg2.items()} if cfg2 else {}),
        **({"config4_clerk_sum": cfg4} if cfg4 else {}), **({"config5_e2e": cfg5} if cfg5 else {}),
        "clerk_combine_x5": {"ms": comb_avg_ms, "share_elements_per_s": n * T * B / (comb_avg_ms * 1e-3),
                             "GBps": comb_bytes / (comb_avg_ms * 1e-3) / 1e9,
                             "frac_of_hbm": comb_bytes / (comb_avg_ms * 1e-3) / 1e9 / peak,
                             "includes": "5 combine launches + library-side NCCL reduce + mod pass" if world > 1 else "5 combine launches"},
    }

    # ---- cpu baseline beside it (bounded sample, rank 0, N=1 only) ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        so, m = oracle_scheme(O)
        cal = cpu_step(O, so, m, 20_000, 1, 1, "cal")
        dim_s = int(max(20_000, min(DIM, args.cpu_seconds / (cal / 20_000))))
        dt = cpu_step(O, so, m, dim_s, 1, 1, "cpu")
        cpu = {"value": dim_s / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"1 participant x {dim_s} secrets (of dim 10M): literal per-batch share-gen + combine, "
                         f"single thread like the reference client; host has {os.cpu_count()} cores"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (Z_p, p=2^61-1): u8 byte limbs on tcgen05 kind::i8 with s32 accumulation, 64-bit integer compose"
                 if "tcgen05" in kernel_name else "u64 (Z_p, p=2^61-1): 32x32->64 IMAD limbs", "data": "synthetic",
        "config": workload_config(args, world), "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "participants_per_step": Te, "steps_timed": e2e_steps, "client_threads": args.e2e_threads, "mode": args.e2e_mode,
                "pageable_value": e2e_pageable, "link": pcie, "cpu_affinity": affinity,
                "link_bound": (None if not pcie else
                               world * Te * dim / max(h2d / (pcie["duplex_each_way_GBps"] * 1e9), d2h / (pcie["duplex_each_way_GBps"] * 1e9))),
                "api": "sda_share_generate per participant + sda_share_combine_rows per clerk (pinned host buffers; one context "
                       "per client thread)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--participants", type=int, default=256, help="resident participants per GPU per step")
    ap.add_argument("--rounds", type=int, default=20, choices=[8, 12, 20], help="ChaCha rounds of the sharing randomness")
    ap.add_argument("--packed-path", default="auto", choices=["auto", "cuda", "tc", "tc1"],
                    help="share-gen kernel: tcgen05 byte-limb GEMM (auto/tc: paired tiles, tc1: first generation) or the "
                         "IMAD.WIDE CUDA-core kernel")
    ap.add_argument("--e2e-participants", type=int, default=4)
    ap.add_argument("--e2e-mode", default="pipelined", choices=["pipelined", "phases"],
                    help="host-buffer leg: clerks of round i-1 concurrently with the participants of round i, or one after the other")
    ap.add_argument("--e2e-threads", type=int, default=6, help="client threads (one context each) of the host-buffer leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-round-sweep", action="store_true", help="skip the ChaCha12/ChaCha8 share-gen launches")
    ap.add_argument("--no-configs45", action="store_true", help="skip the config #4 / config #5 legs (profiling runs)")
    ap.add_argument("--cfg5-participants", type=int, default=1024, help="config #5 participants per GPU (8192 / 8)")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="target CPU seconds per reference step")
    ap.add_argument("--ref-dim", type=int, default=500_000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
