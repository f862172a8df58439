"""Generate tests/golden/*.npz: answers of INDEPENDENT implementations available in this image, frozen as fixtures so
that the oracle (and through it the CUDA path) stays pinned even where those libraries are absent or change:

  meanshift.npz   scikit-learn MeanShift(bandwidth, bin_seeding=True): labels + centres          (sklearn.cluster)
  hdbscan.npz     scikit-learn HDBSCAN(15, 5, eps=0.006, kd_tree): labels                        (sklearn.cluster)
  conv_dense.npz  torch.nn.functional.conv3d / conv_transpose3d on densified toy grids (float64)  (torch)
  ball_query.npz  scipy cKDTree.query_ball_point neighbour sets                                  (scipy.spatial)

    python scripts/make_golden.py        (deterministic: seeded inputs, single-threaded library calls)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
from scipy.spatial import cKDTree
from sklearn.cluster import HDBSCAN, MeanShift

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def blobs(n, D, k, seed, spread=3.0, sigma=0.15, noise=0):
    rng = np.random.default_rng(seed)
    mu = rng.normal(0, spread, (k, D))
    X = mu[rng.integers(0, k, n)] + rng.normal(0, sigma, (n, D))
    if noise:
        X[:noise] = rng.uniform(-2 * spread, 2 * spread, (noise, D))
    return X.astype(np.float32)


# ---- mean shift ----
ms = {}
for i, (n, D, k, seed, h) in enumerate([(600, 5, 7, 0, 0.6), (1500, 3, 12, 1, 0.6), (400, 5, 3, 2, 1.0)]):
    X = blobs(n, D, k, seed)
    m = MeanShift(bandwidth=h, bin_seeding=True).fit(X)
    ms.update({"X%d" % i: X, "h%d" % i: np.float64(h), "labels%d" % i: m.labels_.astype(np.int64),
               "centres%d" % i: m.cluster_centers_.astype(np.float32)})
np.savez_compressed(os.path.join(OUT, "meanshift.npz"), **ms)

# ---- HDBSCAN ----
hd = {}
for i, (n, D, seed) in enumerate([(500, 5, 10), (540, 3, 11), (580, 5, 12)]):
    X = blobs(n, D, 6, seed, noise=n // 20)
    lab = HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006, algorithm="kd_tree").fit_predict(
        X.astype(np.float64))
    hd.update({"X%d" % i: X, "labels%d" % i: lab.astype(np.int64)})
np.savez_compressed(os.path.join(OUT, "hdbscan.npz"), **hd)

# ---- dense convolution ----
G = 8
rng = np.random.default_rng(1)
rows = []
for b in range(2):
    z, y, x = np.nonzero(rng.random((G, G, G)) < 0.35)
    c = np.stack([np.full_like(x, b), x, y, z], 1)
    rows.append(c[rng.permutation(len(c))])
coords = np.concatenate(rows).astype(np.int32)
cin, cout = 16, 32
X = rng.standard_normal((len(coords), cin)).astype(np.float32)
W = (rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32)
ct = torch.as_tensor(coords).long()
dense = torch.zeros(2, cin, G, G, G, dtype=torch.float64)
dense[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]] = torch.as_tensor(X, dtype=torch.float64)
wd = torch.as_tensor(W, dtype=torch.float64).reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)   # [Cout,Cin,kz,ky,kx]
Y = F.conv3d(dense, wd, padding=1)[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]].numpy()
Yt = F.conv_transpose3d(dense, wd.permute(1, 0, 2, 3, 4), padding=1)[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]].numpy()
np.savez_compressed(os.path.join(OUT, "conv_dense.npz"), coords=coords, X=X, W=W, Y=Y, Y_transposed=Yt)

# ---- radius neighbours ----
rng = np.random.default_rng(2)
n = 3000
pos = np.concatenate([rng.normal(c, 0.15, (n // 6, 3)) for c in rng.uniform(-1, 1, (6, 3))]).astype(np.float32)
batch = np.repeat([0, 1], len(pos) // 2)
radius = 0.1
offs, flat = [0], []
for s in (0, 1):
    idx = np.nonzero(batch == s)[0]
    tree = cKDTree(pos[idx].astype(np.float64))
    for q in idx:
        nb = np.sort(idx[tree.query_ball_point(pos[q].astype(np.float64), radius)])
        flat.append(nb)
        offs.append(offs[-1] + len(nb))
np.savez_compressed(os.path.join(OUT, "ball_query.npz"), pos=pos, batch=batch, radius=np.float64(radius),
                    nbr_offsets=np.asarray(offs, np.int64), nbr_flat=np.concatenate(flat).astype(np.int64))
print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})
