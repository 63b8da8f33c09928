"""One `sda_ctx` per thread, several threads at once: `SdaService` is `Send + Sync` in the reference
(protocol/src/methods.rs:13-22), so clients may share a process and call the crypto module concurrently.  The
library keeps no global mutable state outside a context; this drives four contexts from four threads through the
host entry points (pageable and pinned, i.e. plain and sliced paths) and holds every result to the single-thread one."""
import threading

import numpy as np
import pytest

import sda_b200
import util
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import params

pytestmark = pytest.mark.gpu

P61 = params.P61


def participant_job(ctx, tag, dim, pinned):
    """share (packed + additive), mask, combine, reconstruct: everything one client thread would call"""
    rng = np.random.default_rng(tag)
    packed, additive, mask = params.config3(), LSS.Additive(3, P61), LMS.Full(P61)
    sec = util.rand_secrets(rng, dim, P61)
    seed = util.seed_bytes(("threads", tag))
    if pinned:
        h = ctx.pinned_empty(dim)
        h[...] = sec
        out = ctx.pinned_empty(5 * packed.batches(dim)).reshape(5, packed.batches(dim))
        shares = np.array(ctx.share_generate(packed, h, seed, out=out))
    else:
        shares = ctx.share_generate(packed, sec, seed)
    add = ctx.share_generate(additive, sec, seed)
    m, masked = ctx.mask(mask, sec, seed)
    summed = ctx.share_combine(additive, add)
    back = ctx.secret_reconstruct(packed, dim, [(i, shares[i]) for i in range(5)])
    return shares, add, m, masked, summed, back


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_contexts_on_concurrent_threads(pinned):
    dim, nthreads, rounds = 700_001, 4, 3
    solo = sda_b200.Context(0)
    want = {tag: participant_job(solo, tag, dim, pinned) for tag in range(nthreads)}
    for tag in range(nthreads):                      # the job is self-consistent before it is used as a yardstick
        assert np.array_equal(want[tag][5], util.rand_secrets(np.random.default_rng(tag), dim, P61))
    solo.close()

    errors, results = [], {}
    barrier = threading.Barrier(nthreads)

    def worker(tag):
        try:
            ctx = sda_b200.Context(0)
            barrier.wait()
            for _ in range(rounds):
                got = participant_job(ctx, tag, dim, pinned)
                for a, b in zip(got, want[tag]):
                    if not np.array_equal(a, b):
                        raise AssertionError(f"thread {tag}: result differs from the single-thread run")
            results[tag] = True
            ctx.close()
        except Exception as e:                       # noqa: BLE001 - reported below, in the main thread
            errors.append((tag, repr(e)))

    threads = [threading.Thread(target=worker, args=(tag,)) for tag in range(nthreads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    assert len(results) == nthreads
