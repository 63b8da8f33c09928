"""The N>1 path of the clerk sum at world size 2 on CPU (gloo): sharding arithmetic, the single
reduce of canonical partial sums as 64-bit integers, and the final mod on the root.  The per-rank
partial sums come from the oracle here (the test is the checker; on GPUs they come from
sda_share_combine_dev, see bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sda_b200 import multi

P61 = (1 << 61) - 1


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multi.shard_bounds(10, 2, 2)


def test_check_exact():
    multi.check_exact(P61, 8)                      # 8 partials below 2^61 cannot wrap
    multi.check_exact((1 << 63) - 1, 2)
    with pytest.raises(OverflowError):
        multi.check_exact((1 << 63) - 1, 3)
    with pytest.raises(OverflowError):
        multi.check_exact(P61, 9)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, modulus, P, L, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        rng = np.random.default_rng(123)           # same stream on every rank: the full clerk job
        rows = rng.integers(0, modulus, size=(P, L), dtype=np.int64)
        lo, hi = multi.shard_bounds(P, world, rank)
        mine = rows[lo:hi]
        partial = O.canonical(modulus, O.share_combine(modulus, mine)) if hi > lo else np.zeros(L, dtype=np.int64)
        t = torch.from_numpy(np.ascontiguousarray(partial))

        def final_mod(bits):                       # u64 bit patterns -> residues (numpy stands in for the kernel)
            return torch.from_numpy((bits.numpy().view(np.uint64) % np.uint64(modulus)).astype(np.int64))

        got = multi.reduce_partial_sums(t, modulus, dst=0, final_mod=final_mod)
        if rank == 0:
            expect = O.canonical(modulus, O.share_combine(modulus, rows))
            np.save(out_path, np.stack([got.numpy(), expect]))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("modulus,P,L", [(433, 5, 17), (P61, 64, 1000), (P61, 1, 3), ((1 << 62) - 57, 6, 50)])
def test_clerk_sum_reduce_world2(tmp_path, modulus, P, L):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), modulus, P, L, out), nprocs=2, join=True)
    got, expect = np.load(out)
    assert np.array_equal(got, expect)
