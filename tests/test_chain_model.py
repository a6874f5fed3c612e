"""CPU model of the one-launch entropy chain's ALGORITHM (csrc/entropy_partition.cu::entropy_chain_kernel):
linear fine bins, target bins and "pixels in lower bins" from the merged histogram, candidate bands (bin +- 3 delta) probed
through the +-1-bin "near" set, rank inside the candidates = global rank - (lower-bin count - candidates that sit in lower
bins), bit-window select, the early partition of every pixel two or more bins away from the threshold's bins.  The numpy
restatement below follows the kernel step by step; the claim it pins is the kernel's soundness argument: if the fast entropy
is within delta of the exact one, the thresholds equal np.percentile of the EXACT entropies bit for bit (numpy's float32
semantics) and the partition equals the exact comparison for every pixel -- although only a few hundred pixels are ever
evaluated exactly.  (The GPU tests check the kernel itself against the oracle; this runs without a GPU.)"""
import numpy as np
import pytest

BINS, SCALE, DELTA = 4096, np.float32(1024.0), np.float32(1.0e-4)


def fine_bin(h):
    x = h.astype(np.float32) * SCALE
    return np.where(x < 1.0, 0, np.minimum(x.astype(np.int64), BINS - 1)).astype(np.int64)


def fine_lo(b):
    return np.float32(b) * (np.float32(1.0) / SCALE)


def np_rank(n, q):
    """numpy 2.x float32 virtual index (select1 / P2 of the kernel)."""
    nm1 = np.float32(n - 1 if n else 0)
    q32 = np.float32(q) / np.float32(100.0)
    v = np.float32(nm1 * q32)
    fl = np.floor(v)
    if v >= nm1:
        lo = hi = n - 1
    else:
        lo, hi = int(fl), int(fl) + 1
    return lo, hi, np.float32(v - fl)


def lerp(a, b, g):
    d = np.float32(b - a)
    r = np.float32(a + np.float32(d * g))
    if g >= 0.5:
        r = np.float32(b - np.float32(d * np.float32(np.float32(1.0) - g)))
    return r


def chain_model(fast, exact, valid, percents, part_idx):
    n = int(valid.sum())
    fb = fine_bin(fast)
    hist = np.bincount(fb[valid], minlength=BINS)                          # P1
    excl = np.concatenate(([0], np.cumsum(hist)[:-1]))
    incl = np.cumsum(hist)
    granks, gammas = [], []
    for q in percents:                                                     # P2: ranks, bins, lower-bin counts
        lo, hi, g = np_rank(n, q)
        granks += [lo, hi]
        gammas.append(g)
    tbin = [int(np.searchsorted(incl, r, side="right")) for r in granks]
    bands, band_of = [], []
    for t, b in enumerate(tbin):
        if b not in bands:
            bands.append(b)
        band_of.append(bands.index(b))
    stored = fast.copy()
    n_exact = 0
    lists, below = [], []
    for b in bands:
        lo = -np.inf if b == 0 else np.float32(fine_lo(b) - np.float32(3.0) * DELTA)
        hi = np.inf if b == BINS - 1 else np.float32(fine_lo(b + 1) + np.float32(3.0) * DELTA)
        near = valid & (fb + 1 >= b) & (fb <= b + 1)                       # the bitmap probe
        cand = near & (fast >= lo) & (fast < hi)
        stored[cand] = exact[cand]                                         # contract re-evaluation of the candidates only
        n_exact += int(cand.sum())
        binlo = -np.inf if b == 0 else fine_lo(b)
        cand_low = int((cand & (fast < binlo)).sum())                      # candidates that the lower-bin count already holds
        below.append(int(excl[b]) - cand_low)
        lists.append(np.sort(exact[cand]))
    vals = []
    for t, r in enumerate(granks):                                         # P3 (the radix select, as a sort)
        u = band_of[t]
        r0 = r - below[u]
        assert 0 <= r0 < len(lists[u]), "rank-in-candidates invariant violated"
        vals.append(lists[u][r0])
    thr = [lerp(vals[2 * j], vals[2 * j + 1], gammas[j]) for j in range(len(percents))]      # P4
    plo, phi = min(tbin[2 * part_idx], tbin[2 * part_idx + 1]), max(tbin[2 * part_idx], tbin[2 * part_idx + 1])
    drop = np.zeros(fast.shape, bool)
    far_hi = valid & (fb > phi + 1)                                        # decided in P2, before the threshold exists
    far_lo = valid & (fb + 1 < plo)
    mid = valid & ~far_hi & ~far_lo
    drop[far_hi] = True
    drop[mid] = stored[mid] >= thr[part_idx]                               # stored values: exact for every candidate
    return thr, drop, stored, n_exact


@pytest.mark.parametrize("seed,n,frac_ign,percents,part", [
    (0, 200000, 0.0, [90.0, 10.0, 90.0], 0),
    (1, 50000, 0.3, [80.0, 12.5, 87.5], 0),
    (2, 300000, 0.05, [84.0, 16.0, 84.0, 100.0], 2),
    (3, 4097, 0.0, [50.0], 0),
    (4, 100000, 0.0, [99.99, 0.0, 100.0], 1),
    (5, 7, 0.0, [20.0, 80.0], 1),
])
def test_chain_model_thresholds_and_partition_are_exact(seed, n, frac_ign, percents, part):
    rng = np.random.default_rng(seed)
    exact = np.abs(rng.normal(1.4, 0.8, n)).astype(np.float32).clip(0, 3.04)          # entropies of a C = 21 problem
    exact[rng.random(n) < 0.02] = 0.0                                                 # saturated pixels: ties at the bottom
    noise = (rng.random(n).astype(np.float32) - np.float32(0.5)) * np.float32(1.8) * DELTA   # |fast - exact| < 0.9 delta
    fast = (exact + noise).astype(np.float32)
    valid = rng.random(n) >= frac_ign
    thr, drop, stored, n_exact = chain_model(fast, exact, valid, percents, part)
    for j, q in enumerate(percents):
        want = np.percentile(exact[valid], q)
        assert np.float32(thr[j]).view(np.uint32) == np.float32(want).view(np.uint32), (q, thr[j], want)
    assert np.array_equal(drop, valid & (exact >= thr[part]))                          # the reliable / unreliable index set
    for j in range(len(percents)):                                                     # every later comparison (contrastive masks)
        assert np.array_equal(valid & (stored >= thr[j]), valid & (exact >= thr[j]))
        assert np.array_equal(valid & (stored <= thr[j]), valid & (exact <= thr[j]))
    if n >= 50000:
        assert n_exact <= 0.06 * n                                                     # "exact where it matters": a few percent at most


def test_chain_model_massive_ties():
    """Constant entropy (every pixel in one bin, every candidate equal): the invariant and the lerp still hold."""
    n = 5000
    exact = np.full(n, 1.25, np.float32)
    fast = exact + np.float32(3e-6)
    valid = np.ones(n, bool)
    thr, drop, _, n_exact = chain_model(fast, exact, valid, [90.0, 10.0], 0)
    assert thr[0] == np.float32(1.25) and thr[1] == np.float32(1.25)
    assert n_exact == n                                                                # one band holds every pixel
    assert drop.all()                                                                  # entropy >= threshold everywhere
