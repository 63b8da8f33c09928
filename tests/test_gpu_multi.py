"""The multi-GPU clerk sum behind the C ABI (sda_ctx_create_multi, sda_ctx_comm_init_rank, sda_partial_sums_reduce_dev,
sda_share_combine_ranks_dev, sda_share_combine_multi_dev, sda_share_combine_rows_multi): NCCL is called by the library,
no torch.distributed in the data path.  One-GPU boxes run the degenerate communicator (1 rank / 1 device); the tests
that need two devices skip themselves there and are run with `gpurun --gpus 2`."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import sda_b200
from sda_b200 import _lib, params
from sda_b200 import LinearSecretSharingScheme as LSS

pytestmark = pytest.mark.gpu
P61 = (1 << 61) - 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def create_multi(devices):
    lib = sda_b200.load()
    h = C.c_void_p()
    arr = (C.c_int * len(devices))(*devices)
    rc = lib.sda_ctx_create_multi(arr, len(devices), C.byref(h))
    assert rc == 0, lib.sda_last_error(None).decode()
    return lib, h


@pytest.mark.parametrize("modulus", [433, P61, (1 << 62) - 57])
def test_single_rank_communicator_is_the_plain_combine(ctx, oracle, modulus):
    """world size 1: comm_init_rank succeeds, the reduce is the identity, combine_ranks == combine"""
    c = sda_b200.Context(0)
    c.comm_init_rank(sda_b200.Context.nccl_unique_id(), 1, 0)
    assert (c.comm_rank(), c.comm_size()) == (0, 1)
    rng = np.random.default_rng(5)
    P, L = 37, 1001
    rows = rng.integers(0, modulus, size=(P, L), dtype=np.int64)
    d_rows = torch.from_numpy(rows).cuda()
    d_out = torch.empty(L, dtype=torch.int64, device="cuda")
    s = LSS.Additive(3, modulus)
    c.share_combine_ranks_dev(s, d_rows, L, P, L, d_out)
    c.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), oracle.canonical(modulus, oracle.share_combine(modulus, rows)))
    c.close()


def test_create_multi_rejects_bad_device_lists():
    lib = sda_b200.load()
    h = C.c_void_p()
    assert lib.sda_ctx_create_multi((C.c_int * 2)(0, 0), 2, C.byref(h)) == _lib.SDA_ERR_INVALID
    assert b"twice" in lib.sda_last_error(None)
    assert lib.sda_ctx_create_multi((C.c_int * 1)(0), 0, C.byref(h)) == _lib.SDA_ERR_INVALID
    assert lib.sda_ctx_create_multi((C.c_int * 1)(99), 1, C.byref(h)) != 0


def test_one_device_group_combines_host_rows(oracle):
    lib, h = create_multi([0])
    assert lib.sda_ctx_multi_count(h) == 1 and lib.sda_ctx_multi_member(h, 0) == h.value and not lib.sda_ctx_multi_member(h, 1)
    s = params.config4()
    rng = np.random.default_rng(6)
    rows = [rng.integers(0, P61, size=513, dtype=np.int64) for _ in range(9)]
    ptrs = (C.c_void_p * 9)(*[r.ctypes.data for r in rows])
    lens = (C.c_size_t * 9)(*[513] * 9)
    out = np.empty(513, dtype=np.int64)
    n = C.c_size_t(0)
    assert lib.sda_share_combine_rows_multi(h, C.byref(s.c), ptrs, lens, 9, out.ctypes.data_as(C.c_void_p), C.byref(n)) == 0
    assert n.value == 513 and np.array_equal(out, oracle.canonical(P61, oracle.share_combine(P61, np.stack(rows))))
    lens[4] = 512
    assert lib.sda_share_combine_rows_multi(h, C.byref(s.c), ptrs, lens, 9, out.ctypes.data_as(C.c_void_p), C.byref(n)) == _lib.SDA_ERR_INVALID
    assert lib.sda_last_error(h) == b"Wrong dimension"
    lib.sda_ctx_destroy(h)


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")


@needs2
@pytest.mark.parametrize("modulus,P,L", [(P61, 301, 100_003), (433, 7, 33), ((1 << 62) - 57, 64, 5000), (P61, 1, 9)])
def test_two_device_group_host_rows_and_device_rows(oracle, modulus, P, L):
    """one process, two GPUs: rows sharded over the devices, one NCCL reduce (or gather + fold above 2^63 / ranks)"""
    lib, h = create_multi([0, 1])
    assert lib.sda_ctx_multi_count(h) == 2
    s = LSS.Additive(3, modulus)
    rng = np.random.default_rng(P * 31 + L)
    rows = rng.integers(0, modulus, size=(P, L), dtype=np.int64)
    expect = oracle.canonical(modulus, oracle.share_combine(modulus, rows))
    # host rows (what the Rust clerk passes)
    ptrs = (C.c_void_p * P)(*[rows[i].ctypes.data for i in range(P)])
    lens = (C.c_size_t * P)(*[L] * P)
    out = np.empty(L, dtype=np.int64)
    n = C.c_size_t(0)
    assert lib.sda_share_combine_rows_multi(h, C.byref(s.c), ptrs, lens, P, out.ctypes.data_as(C.c_void_p), C.byref(n)) == 0, \
        lib.sda_last_error(h)
    assert np.array_equal(out, expect)
    # rows already resident on the two devices, uneven split (device 1 may get nothing)
    cut = (2 * P) // 3
    d0 = torch.from_numpy(rows[:cut]).to("cuda:0")
    d1 = torch.from_numpy(rows[cut:]).to("cuda:1") if cut < P else torch.empty((0, L), dtype=torch.int64, device="cuda:1")
    part1 = torch.empty(L, dtype=torch.int64, device="cuda:1")
    d_out = torch.empty(L, dtype=torch.int64, device="cuda:0")
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    shares = (C.c_void_p * 2)(d0.data_ptr(), d1.data_ptr())
    parts = (C.c_void_p * 2)(None, part1.data_ptr())
    per = (C.c_size_t * 2)(cut, P - cut)
    assert lib.sda_share_combine_multi_dev(h, C.byref(s.c), shares, L, per, L, parts, C.c_void_p(d_out.data_ptr())) == 0, \
        lib.sda_last_error(h)
    assert lib.sda_ctx_synchronize(h) == 0
    assert np.array_equal(d_out.cpu().numpy(), expect)
    lib.sda_ctx_destroy(h)


WORKER = r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, {root!r})
import sda_b200
from sda_b200 import LinearSecretSharingScheme as LSS
from oracle import oracle as O
rank, world, idfile, modulus = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
torch.cuda.set_device(rank)
ctx = sda_b200.Context(rank)
if rank == 0:
    open(idfile + ".tmp", "wb").write(sda_b200.Context.nccl_unique_id()); os.rename(idfile + ".tmp", idfile)
import time
while not os.path.exists(idfile): time.sleep(0.01)
ctx.comm_init_rank(open(idfile, "rb").read(), world, rank)
P, L = 50, 20011
rows = np.random.default_rng(9).integers(0, modulus, size=(P, L), dtype=np.int64)      # the whole clerk job, same on every rank
lo, hi = rank * P // world, (rank + 1) * P // world
d_rows = torch.from_numpy(np.ascontiguousarray(rows[lo:hi])).cuda()
d_out = torch.empty(L, dtype=torch.int64, device="cuda")
ctx.share_combine_ranks_dev(LSS.Additive(3, modulus), d_rows, L, hi - lo, L, d_out, root=0)
ctx.synchronize()
if rank == 0:
    O.build()
    assert np.array_equal(d_out.cpu().numpy(), O.canonical(modulus, O.share_combine(modulus, rows)))
    print("rank0 ok")
"""


@needs2
@pytest.mark.parametrize("modulus", [P61, (1 << 62) - 57])
def test_two_ranks_two_processes(tmp_path, modulus):
    """one process per GPU: the id travels through a file, the collective is sda_share_combine_ranks_dev"""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    idfile = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, str(modulus)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "rank0 ok" in outs[0]
