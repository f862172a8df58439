"""oracle/transforms_ref.py against an independent dictionary-based definition of "one row per voxel"."""
import numpy as np

from oracle import transforms_ref as tr


def test_grid_sample_against_dictionary_definition():
    rng = np.random.default_rng(0)
    pos = rng.uniform(-3, 3, (5000, 3)).astype(np.float32)
    batch = rng.integers(0, 3, 5000)
    idx, cluster, coords = tr.grid_sample(pos, 0.25, batch)
    last = {}
    for i in range(len(pos)):
        key = (int(batch[i]),) + tuple(int(v) for v in np.round(pos[i] / np.float32(0.25))[::-1])   # (b, z, y, x)
        last[key] = i
    want = [last[k] for k in sorted(last)]                 # ascending voxel id == lexicographic (b, z, y, x)
    assert idx.tolist() == want
    assert len(np.unique(cluster)) == len(want) and np.array_equal(np.sort(np.unique(cluster)), np.arange(len(want)))
    for v in range(0, len(want), 97):                      # every member of voxel v quantises to the same cell
        m = cluster == v
        assert len(np.unique(coords[m], axis=0)) == 1 and len(np.unique(batch[m])) == 1 and idx[v] == np.nonzero(m)[0].max()


def test_cylinder_is_a_disc_in_xy():
    rng = np.random.default_rng(1)
    pos = rng.uniform(-10, 10, (4000, 3)).astype(np.float32)
    ind = tr.cylinder(pos, (1.0, -2.0), 4.0)
    r = np.hypot(pos[:, 0].astype(np.float64) - 1.0, pos[:, 1].astype(np.float64) + 2.0)
    assert np.array_equal(ind, np.nonzero(r <= 4.0)[0])
