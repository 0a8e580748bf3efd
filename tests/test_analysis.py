"""the consumer side of the .bin (voxelmotionviewer.py:28-96).
CPU (`-m "not gpu"`): the host mirror of numpy's percentile arithmetic against np.percentile itself, and the golden
vectors of the reference's own functions re-derived with numpy (pins the golden file).  GPU: radix-select order
statistics and ordered extraction against the golden vectors, against numpy on random grids, and at 512^3."""
import numpy as np
import pytest
import torch

from conftest import bits, golden

from pixeltovoxelprojector_b200 import analysis

CASES = [("sparse", i) for i in range(3)] + [("dense", i) for i in range(4)] + [("empty", 0)] + [("inexact", i) for i in range(3)]


def _hard(g, name, i):
    """hard_thresh as the golden run passed it: a Python number (weak: compared in float32) or np.float64 (compared in double); None = percentile."""
    h = float(g[f"{name}_{i}_hard"])
    if h < 0:
        return None
    if bool(g[f"{name}_{i}_hard_is_np64"]):
        return np.float64(h)
    return int(h) if h == int(h) else h


def test_percentile_index_arithmetic_equals_numpy():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 100000))
        a = (rng.integers(0, 40, n) * rng.choice([0, 1, 1.5])).astype(np.float32) if trial % 2 else rng.normal(size=n).astype(np.float32)
        q = float(rng.choice([0, 100, 50, 99.5, 99.9, 12.3456, rng.uniform(0, 100)]))
        want = np.percentile(a, q)
        s = np.sort(a)
        p, nx, g = analysis.percentile_indices(n, q, np.float32)
        got = analysis.lerp_like_numpy(s[p], s[nx], g)
        assert type(got) is type(want) and got.tobytes() == want.tobytes(), (n, q)
    with pytest.raises(ValueError):
        analysis.percentile_indices(10, 101.0)
    # 512^3 cells: numpy forms the virtual index in float32 (the array's dtype) -- the mirror must do the same
    p, nx, g = analysis.percentile_indices(512 ** 3, 99.5)
    q32 = np.true_divide(99.5, np.float32(100))
    assert p == int(np.floor((512 ** 3 - 1) * q32)) and g.dtype == np.float32
    assert nx == p      # float32 spacing at 1.3e8 is 8: numpy's `previous + 1` rounds back to `previous` (mirrored, not fixed)


@pytest.mark.parametrize("name,i", CASES)
def test_golden_is_the_numpy_recipe(name, i):
    """The committed golden vectors (made by the reference's functions) equal the recipe of :56-96 re-evaluated here."""
    g = golden("viewer_golden.npz")
    grid, vs, center = g[f"{name}_grid"], g[f"{name}_voxel_size"], g[f"{name}_center"]
    hard = _hard(g, name, i)
    thresh = hard if hard is not None else np.percentile(grid.ravel(), float(g[f"{name}_{i}_percentile"]))
    coords = np.argwhere(grid > thresh)
    assert len(coords) == len(g[f"{name}_{i}_points"]) and bool(g[f"{name}_{i}_none"]) == (len(coords) == 0)
    assert np.array_equal(grid[coords[:, 0], coords[:, 1], coords[:, 2]], g[f"{name}_{i}_intensities"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", CASES)
def test_extract_top_percentile_matches_reference_golden(ctx, name, i, capsys, tmp_path):
    g = golden("viewer_golden.npz")
    import pixeltovoxelprojector_b200 as pvp
    path = str(tmp_path / "g.bin")
    pvp.write_bin(path, g[f"{name}_grid"], float(g[f"{name}_voxel_size"]))
    grid, vs = pvp.load_voxel_grid(path)                       # CUDA tensor + np.float32, like the reference's pair
    assert grid.is_cuda and isinstance(vs, np.float32) and vs == g[f"{name}_voxel_size"]
    hard = _hard(g, name, i)
    kw = dict(use_hard_thresh=True, hard_thresh=hard) if hard is not None else dict(percentile=float(g[f"{name}_{i}_percentile"]))
    pts, inten = pvp.extract_top_percentile_z_up(grid, vs, g[f"{name}_center"], **kw)
    if bool(g[f"{name}_{i}_none"]):
        assert pts is None and inten is None
        assert "No voxels above threshold" in capsys.readouterr().out
        return
    assert pts.dtype == torch.float64 and inten.dtype == torch.float32
    assert np.array_equal(bits(pts.cpu().numpy()), bits(g[f"{name}_{i}_points"]))             # bit-exact, same order
    assert np.array_equal(bits(inten.cpu().numpy()), bits(g[f"{name}_{i}_intensities"]))


@pytest.mark.gpu
def test_order_statistics_and_argwhere_random(ctx):
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(3, 70))
        kind = trial % 4
        if kind == 0:
            a = rng.normal(size=(n, n, n)).astype(np.float32)                       # negatives, no ties
        elif kind == 1:
            a = rng.integers(0, 4, (n, n, n)).astype(np.float32)                    # heavy ties
        elif kind == 2:
            a = np.where(rng.random((n, n, n)) < 0.02, rng.integers(1, 1 << 20, (n, n, n)), 0).astype(np.float32)
        else:
            a = (rng.standard_cauchy((n, n, n)) * 1e3).astype(np.float32)
            a.flat[:3] = [-0.0, 0.0, np.inf]
        t = torch.as_tensor(a, device="cuda")
        for q in (0.0, 100.0, 50.0, 99.5, float(rng.uniform(0, 100))):
            want = np.percentile(a.ravel(), q)
            got = analysis.grid_percentile(t, q)
            assert type(got) is np.float32 and got.tobytes() == want.tobytes(), (trial, q, want, got)
        thr = np.percentile(a.ravel(), 90.0)
        assert np.array_equal(analysis.argwhere_above(t, thr).cpu().numpy(), np.argwhere(a > thr))
        assert analysis.argwhere_above(t, np.inf).shape == (0, 3)


@pytest.mark.gpu
def test_full_size_512_properties(ctx):
    """512^3 (134 M cells): the percentile is an order statistic (counts below/above bracket the rank), the extraction is
    ordered, complete and consistent with torch's own comparison."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    n = 512
    grid = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    hits = torch.randint(0, n ** 3, (3_000_000,), device="cuda", generator=gen)
    grid.view(-1).index_add_(0, hits, torch.randint(3, 256, (3_000_000,), device="cuda", generator=gen).float())
    thr = analysis.grid_percentile(grid, 99.5)
    assert thr.tobytes() == np.percentile(grid.cpu().numpy().ravel(), 99.5).tobytes()   # numpy itself on the same cells
    p, nx, gamma = analysis.percentile_indices(n ** 3, 99.5)
    below, not_above = int((grid < float(thr)).sum()), int((grid <= float(thr)).sum())
    assert below <= nx and not_above > p                                # thr lies between the two order statistics
    pts, inten = analysis.extract_top_percentile_z_up(grid, np.float32(2.0), np.array([0, 0, 500]), percentile=99.5)
    mask = grid > float(thr)
    assert len(inten) == int(mask.sum()) > 1000
    assert torch.equal(inten, grid[mask])                               # C order == torch's masked-select order
    idx = torch.nonzero(mask)
    want = torch.stack([-512.0 + (idx[:, 2].double() + 0.5) * 2.0, -512.0 + (idx[:, 1].double() + 0.5) * 2.0,
                        500.0 - 512.0 + (idx[:, 0].double() + 0.5) * 2.0], 1)
    assert torch.equal(pts, want)
