"""Host-view Cholesky in arrival waves (api.cu: cholesky_host, potrf.cu: cholesky_lower_wave; option chol_waves): block columns cross
PCIe in order on their own stream while the columns that have arrived are factored -- a large-K catch-up product per wave, then a
right-looking sweep restricted to the wave's columns.  The result must be the factor of cholesky.rs:51-83 whatever the number of
waves, the host layout (row-major / column-major), the memory kind (pageable staging / pinned) and the variant (dirty / clean);
a non-positive pivot inside a later wave must surface with its global index (cholesky.rs:69-71)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 8192
EPS = 2.220446049250313e-16


@pytest.fixture(scope="module")
def L():
    import linfa_linalg_b200 as L
    L.engine()
    return L


@pytest.fixture(scope="module")
def spd():
    rng = np.random.default_rng(8192)
    s = rng.uniform(-1, 1, (N, N))
    s = (s + s.T) * 0.5
    s[np.diag_indices(N)] += N
    return s


def _factor(L, a, waves, clean=False):
    e = L.Engine(0)
    e.set_option("chol_waves", waves)
    return (L.cholesky_inplace if clean else L.cholesky_inplace_dirty)(a, eng=e)


def _residual(a0, l):
    import torch
    lt = torch.tril(torch.from_numpy(l).cuda())
    r = lt @ lt.t() - torch.from_numpy(a0).cuda()
    return float(torch.tril(r).norm() / np.linalg.norm(a0))


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("waves", [2, 3, 4])
def test_waves_match_single_pass(L, spd, order, waves):
    a1 = np.array(spd, order=order)
    _factor(L, a1, 1)
    aw = np.array(spd, order=order)
    _factor(L, aw, waves)
    assert _residual(spd, aw) <= 64 * EPS
    assert np.max(np.abs(np.tril(aw) - np.tril(a1))) <= 1e3 * EPS * N          # same factor up to the summation order of the updates
    iu = np.triu_indices(N, 1)
    assert np.array_equal(aw[iu], spd[iu])                                       # dirty variant: the strict upper triangle is the caller's


def test_waves_clean_variant(L, spd):
    a = np.array(spd)
    _factor(L, a, 3, clean=True)
    assert _residual(spd, a) <= 64 * EPS
    assert not np.any(np.triu(a, 1))


def test_waves_pinned_host(L, spd):
    import torch
    t = torch.from_numpy(spd.copy()).pin_memory()
    a = t.numpy()
    _factor(L, a, 3)
    assert _residual(spd, a) <= 64 * EPS


@pytest.mark.parametrize("bad", [700, 3000, 6000])
def test_waves_report_global_failure_index(L, spd, bad):
    a = np.array(spd)
    a[bad, bad] = -1.0
    with pytest.raises(L.NotPositiveDefinite) as ei:
        _factor(L, a, 3)
    assert ei.value.index == bad
    a = np.array(spd)
    a[bad, bad] = -1.0
    with pytest.raises(L.NotPositiveDefinite) as e1:
        _factor(L, a, 1)
    assert e1.value.index == bad
