"""One box, N GPUs: how the hot path shards.

Share generation and masking are independent per participant: participants are dealt to ranks in
contiguous blocks and no collective is involved.  The clerk sum (`ShareCombiner::combine`,
client/src/crypto/sharing/combiner.rs:15-29) is the one step with an exchange: every rank folds its
own participants into canonical partial sums in [0, m), the partials are added as 64-bit integers
with a single `reduce(SUM)` (NCCL over NVLink on GPUs), and the root applies one final `mod m`
(`sda_mod_reduce_u64_dev`).  The integer sum is exact while world_size * (m - 1) < 2^64, i.e. for
8 ranks at any 61-bit modulus; larger products are refused rather than wrapped.

This module holds only the host-side plumbing (sharding arithmetic and the collective); the
arithmetic itself is the library's.  It is exercised on CPU by tests/test_multi_gloo.py with the
gloo backend at world size 2, and on GPUs by bench.py under torchrun.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous block [lo, hi) of `total` participants owned by `rank`; blocks differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def check_exact(modulus, world):
    """The uint64 sum of `world` canonical partials must not wrap."""
    if modulus <= 0:
        raise ValueError("modulus must be positive")
    if world * (modulus - 1) >= 1 << 64:
        raise OverflowError(f"{world} partial sums below {modulus} can exceed 2^64: reduce in two levels")


def reduce_partial_sums(partial, modulus, dst=0, group=None, final_mod=None):
    """Sum canonical per-rank partial clerk sums onto `dst` and reduce them mod `modulus`.

    partial    int64 tensor (any shape) holding residues in [0, modulus): this rank's combine output.
               Overwritten with the 64-bit integer sum on `dst` (bit pattern = u64).
    final_mod  callable(tensor_u64_bits) -> tensor applying `x mod modulus` to the u64 bit patterns;
               on GPUs `lambda t: ctx.mod_reduce_dev(modulus, t, t.numel(), out, unsigned=True)`.
    Returns final_mod's result on `dst`, None elsewhere.  One collective, no other traffic.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    check_exact(modulus, world)
    if partial.dtype != torch.int64:
        raise TypeError("partial sums are int64 (i64 shares)")
    if world > 1:
        dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM, group=group)   # two's-complement add == u64 add
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank != dst:
        return None
    return final_mod(partial) if final_mod is not None else partial
