"""CPU restatement of sklearn.cluster.MeanShift(bandwidth, bin_seeding=True) -- TEST INFRASTRUCTURE ONLY (the product
path is csrc/meanshift.cu behind panopticsegforlargescalepointcloud_b200/meanshift.py; nothing outside tests/ may import
this file).

Follows scikit-learn's own source (the reference imports it: torch_points3d/utils/meanshift_cluster.py:4,9-18; pinned
0.24.2 in poetry.lock:2028-2029, 1.9 in this image -- same algorithm):
    sklearn/cluster/_mean_shift.py:247-297  get_bin_seeds
    sklearn/cluster/_mean_shift.py:108-132  _mean_shift_single_seed
    sklearn/cluster/_mean_shift.py:470-560  MeanShift.fit (duplicate removal, nearest-centre labels)
with the KD-tree radius / nearest queries replaced by brute force in float64 (a KD-tree returns the same sets).
PINNED against sklearn.cluster.MeanShift itself in tests/test_oracle_meanshift.py (labels and centres)."""
import numpy as np


def bin_seeds(X, bin_size, min_bin_freq=1):
    bins = {}
    for p in X:
        key = tuple(np.round(p / bin_size))
        bins[key] = bins.get(key, 0) + 1
    seeds = np.array([k for k, f in bins.items() if f >= min_bin_freq], dtype=np.float32)
    if len(seeds) == len(X):
        return X
    return seeds * bin_size


def single_seed(mean, X64, X, bandwidth, max_iter):
    stop = 1e-3 * bandwidth
    it = 0
    while True:
        d2 = ((X64 - np.asarray(mean, dtype=np.float64)) ** 2).sum(1)
        inside = np.nonzero(d2 <= bandwidth * bandwidth)[0]
        if len(inside) == 0:
            break
        old = mean
        mean = np.mean(X[inside], axis=0)
        if np.linalg.norm(mean - old) <= stop or it == max_iter:
            break
        it += 1
    return tuple(mean), len(inside), it


def mean_shift(X, bandwidth, bin_seeding=True, min_bin_freq=1, max_iter=300):
    """-> (labels int64 [n], centres float [c, D]) exactly as MeanShift(...).fit(X).labels_ / .cluster_centers_"""
    X = np.ascontiguousarray(X)
    X64 = X.astype(np.float64)
    seeds = bin_seeds(X, bandwidth, min_bin_freq) if bin_seeding else X
    centre_intensity = {}
    for s in seeds:
        c, k, _ = single_seed(s, X64, X, bandwidth, max_iter)
        if k:
            centre_intensity[c] = k
    if not centre_intensity:
        raise ValueError("No point was within bandwidth of any seed")
    srt = sorted(centre_intensity.items(), key=lambda t: (t[1], t[0]), reverse=True)
    centres = np.array([t[0] for t in srt])
    c64 = centres.astype(np.float64)
    unique = np.ones(len(centres), dtype=bool)
    for i in range(len(centres)):
        if unique[i]:
            unique[((c64 - c64[i]) ** 2).sum(1) <= bandwidth * bandwidth] = False
            unique[i] = True
    centres = centres[unique]
    d2 = ((X64[:, None, :] - centres.astype(np.float64)[None, :, :]) ** 2).sum(2)
    return d2.argmin(1).astype(np.int64), centres


def blobs(n, D, n_inst, seed, spread=3.0, sigma=0.15):
    """Synthetic embeddings of SURVEY 8d: mu_inst ~ N(0, spread^2 I), points mu + N(0, sigma^2 I), float32."""
    rng = np.random.default_rng(seed)
    mu = rng.normal(0, spread, (n_inst, D))
    inst = rng.integers(0, n_inst, n)
    return (mu[inst] + rng.normal(0, sigma, (n, D))).astype(np.float32), inst
