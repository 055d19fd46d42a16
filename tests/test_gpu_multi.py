"""Multi-GPU behind the C ABI (csrc/multi.cu, include/linfa_b200.h `lfb_*_multi_*`): ONE process drives several devices
through ctypes, the way the Rust shim would.  Tall-skinny qr_into is row-sharded (ncclAllGather of the R factors +
ncclBroadcast of U' / diag inside the library), the batched factorisations are batch-sharded.  Parity bar: the reference's
own compact factor from the CPU oracle, elementwise (qr.rs:29-45), exactly like the single-device tests.

World sizes: every size from 1 to the number of visible devices (capped at 8).  On a one-GPU box only the G = 1 cases run
(no collective exists there); `gpurun --gpus 2` (profiles/r2_multi_gpu.txt) and the 8-GPU bench exercise G >= 2."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
EPS = 2.220446049250313e-16


def _ndev():
    import torch
    return min(torch.cuda.device_count(), 8)


def _worlds():
    import os
    try:
        n = _ndev()
    except Exception:
        n = 1
    only = os.environ.get("LFB_TEST_WORLDS")          # e.g. "2,8": an 8-GPU box is charged 8x, so do not sweep 1..8 there
    if only:
        return [int(w) for w in only.split(",") if 1 <= int(w) <= max(n, 1)] or [1]
    return list(range(1, max(n, 1) + 1))


@pytest.fixture(scope="module")
def multis():
    from linfa_linalg_b200.dist import MultiEngine
    made = {}

    def get(g):
        if g not in made:
            made[g] = MultiEngine(n_devices=g)
            made[g].set_option("tsqr_chunk", 1024)
        return made[g]
    yield get
    for m in made.values():
        m.close()


@pytest.mark.parametrize("world", _worlds())
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("rows,cols,order", [(5000, 24, "C"), (9001, 64, "F"), (40000, 130, "C")])
def test_qr_tsqr_multi_matches_reference_compact_factor(multis, world, rows, cols, order, dt):
    m = multis(world)
    assert m.nccl_ranks == (world if world > 1 else 0)
    eps = EPS if dt == np.float64 else 1.1920929e-07
    a0 = np.random.default_rng(rows + cols + world).uniform(-1, 1, (rows, cols)).astype(dt)
    ref = np.asfortranarray(a0); dref = O.qr(ref)
    a = np.array(a0, order=order)
    l0 = m.launch_count
    diag = m.qr_tsqr_into(a)
    assert m.launch_count > l0
    t = 64 * cols * eps * np.linalg.norm(a0.astype(np.float64), 2)
    assert np.max(np.abs(np.tril(a) - np.tril(ref))) <= 64 * cols * eps * 4           # unit-norm reflectors: absolute
    assert np.max(np.abs(np.triu(a[:cols], 1) - np.triu(ref[:cols], 1))) <= t
    assert np.max(np.abs(diag - dref)) <= t
    assert np.array_equal(np.signbit(diag), np.signbit(dref))
    # the reference's consumers accept it (tests/qr.rs:20-27)
    q = O.generate_q(a.copy(), diag)
    r = O.qr_into_r(a.copy(), diag)
    assert np.linalg.norm(q.astype(np.float64) @ r - a0) <= 64 * cols * eps * np.linalg.norm(a0.astype(np.float64))


@pytest.mark.parametrize("world", _worlds())
def test_tsqr_r_multi(multis, world):
    m = multis(world)
    rows, cols = 30011, 96
    a0 = np.random.default_rng(7 + world).uniform(-100, 100, (rows, cols))
    keep = a0.copy()
    r = m.tsqr_r(a0)
    np.testing.assert_array_equal(a0, keep)                       # read only
    ref = np.asfortranarray(keep); dref = O.qr(ref); r_ref = O.qr_into_r(ref, dref)
    assert np.all(np.diag(r) >= 0) and np.all(np.tril(r, -1) == 0)
    assert np.max(np.abs(r - r_ref)) <= 64 * cols * EPS * np.linalg.norm(keep, 2) * 8


@pytest.mark.parametrize("world", _worlds())
def test_short_matrix_and_errors(multis, world):
    m = multis(world)
    a0 = np.random.default_rng(3).uniform(-1, 1, (70, 40))        # fewer than world*cols rows for world >= 2: device 0 alone
    ref = a0.copy(); dref = O.qr(ref)
    a = a0.copy()
    diag = m.qr_tsqr_into(a)
    t = 64 * 70 * EPS * np.linalg.norm(a0)
    assert np.max(np.abs(a - ref)) <= t and np.max(np.abs(diag - dref)) <= t
    with pytest.raises(RuntimeError, match="status 2"):           # NotThin, qr.rs:34-36
        m.qr_tsqr_into(np.zeros((3, 5)))
    assert m.qr_tsqr_into(np.zeros((0, 0))).shape == (0,)         # qr.rs:383-388


@pytest.mark.parametrize("world", _worlds())
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_batched_multi(multis, world, dt):
    m = multis(world)
    eps = EPS if dt == np.float64 else 1.1920929e-07
    batch = 4099                                                   # not divisible by the world size
    a0 = np.random.default_rng(5).uniform(-1, 1, (batch, 32, 32)).astype(dt)
    a0[7, :, 3] = 0
    ref = a0.copy(); dref = O.qr_batched(ref)
    a = a0.copy(); d = m.qr_batched(a)
    t = 16 * 32 * eps * np.sqrt(32)
    assert np.max(np.abs(a - ref)) <= t and np.max(np.abs(d - dref)) <= t
    assert np.array_equal(np.signbit(d), np.signbit(dref))
    g = np.random.default_rng(6).uniform(-1, 1, (batch, 20, 20))
    s0 = (g @ g.transpose(0, 2, 1) + 20 * np.eye(20)[None]).astype(dt)
    refc = s0.copy(); assert O.cholesky_batched(refc, True) == (-1, -1)
    s = s0.copy(); assert m.cholesky_batched(s, True) == (-1, -1)
    assert np.max(np.abs(s - refc)) <= 16 * 20 * eps * np.sqrt(20) * np.max(np.abs(s0))
    bad = s0.copy()
    bad[batch - 2, 19, 19] = -1.0                                  # fails on the LAST device ...
    bad[batch // 2, 4, 4] = -2.0                                   # ... but this one comes first in batch order
    assert m.cholesky_batched(bad, False) == (batch // 2, 4)


@pytest.mark.parametrize("world", _worlds())
def test_device_resident_blocks(multis, world):
    """lfb_qr_tsqr_multi_dev_f64 / lfb_tsqr_r_multi_dev_f64 on per-device torch tensors; uneven row blocks."""
    import torch
    m = multis(world)
    n = 48
    rows = [6000 + 37 * i for i in range(world)]
    rng = np.random.default_rng(11)
    parts = [rng.uniform(-1, 1, (r, n)) for r in rows]
    full = np.vstack(parts)
    ref = np.asfortranarray(full); dref = O.qr(ref); r_ref = O.qr_into_r(ref, dref)
    blocks = [torch.from_numpy(np.ascontiguousarray(p.T)).to(f"cuda:{i}") for i, p in enumerate(parts)]
    diags = [torch.zeros(n, dtype=torch.float64, device=f"cuda:{i}") for i in range(world)]
    rs = [torch.zeros((n, n), dtype=torch.float64, device=f"cuda:{i}") for i in range(world)]
    for i in range(world):
        torch.cuda.synchronize(i)
    m.time_begin()
    m.qr_tsqr_dev(blocks, diags, rs)
    ms = m.time_end()
    assert ms > 0
    t = 64 * n * EPS * np.linalg.norm(full, 2)
    got = np.vstack([b.cpu().numpy().T for b in blocks])
    assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= 64 * n * EPS * 4
    assert np.max(np.abs(np.triu(got[:n], 1) - np.triu(ref[:n], 1))) <= t
    for i in range(world):
        assert np.max(np.abs(diags[i].cpu().numpy() - dref)) <= t
        assert np.max(np.abs(rs[i].cpu().numpy().T - r_ref)) <= t
    blocks = [torch.from_numpy(np.ascontiguousarray(p.T)).to(f"cuda:{i}") for i, p in enumerate(parts)]
    for i in range(world):
        torch.cuda.synchronize(i)
    m.tsqr_r_dev(blocks, rs)
    m.synchronize()
    for i in range(world):
        assert np.max(np.abs(rs[i].cpu().numpy().T - r_ref)) <= t * 8
