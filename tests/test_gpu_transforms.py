"""Device GridSampling3D / CylinderSampling against the CPU oracle (bit-exact indices and coordinates)."""
import numpy as np
import pytest
import torch

from oracle import transforms_ref as tr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,scale", [(0.12, 1.0), (0.6, 1.0), (0.12, 900.0)])
def test_grid_sampling_indices_and_attributes(cuda_device, size, scale):
    """size 0.12: mostly one point per voxel; 0.6: ~8 points per voxel (the hash de-duplication does the work);
    scale 900: a grid wider than 65535 cells per axis -> the sort-all-points fallback.  Same contract in all three."""
    from panopticsegforlargescalepointcloud_b200 import transforms as T
    rng = np.random.default_rng(2)
    n = 60000
    pos = (rng.uniform(-8, 8, (n, 3)) * scale).astype(np.float32)
    batch = np.sort(rng.integers(0, 2, n))
    y = rng.integers(0, 9, n)
    perm = rng.permutation(n)
    want_idx, want_cluster, want_coords = tr.grid_sample(pos[perm], size, batch[perm])
    d = {"pos": torch.from_numpy(pos).to(cuda_device), "batch": torch.from_numpy(batch).to(cuda_device),
         "y": torch.from_numpy(y).to(cuda_device), "meta": torch.zeros(3, device=cuda_device)}
    out = T.GridSampling3D(size, quantize_coords=True, mode="last", return_inverse=True)(d, perm=torch.from_numpy(perm).to(cuda_device))
    assert (len(want_idx) < 0.6 * n) == (size == 0.6)         # 0.6: about two points per occupied voxel
    assert np.array_equal(out["pos"].cpu().numpy(), pos[perm][want_idx])
    assert np.array_equal(out["y"].cpu().numpy(), y[perm][want_idx])
    assert np.array_equal(out["batch"].cpu().numpy(), batch[perm][want_idx])
    assert out["coords"].dtype == torch.int32 and np.array_equal(out["coords"].cpu().numpy(), want_coords[want_idx].astype(np.int32))
    assert np.array_equal(out["inverse_indices"].cpu().numpy(), want_cluster)
    assert out["meta"].shape[0] == 3 and float(out["grid_size"][0]) == pytest.approx(size)
    # the result feeds the hot path: one row per voxel, so the coordinate map accepts it
    from panopticsegforlargescalepointcloud_b200 import me
    if scale == 1.0:
        c4 = torch.cat([out["batch"].int().unsqueeze(1), out["coords"]], 1)
        assert me.CoordinateManager(c4).get_map(1).n == c4.shape[0]


def test_cylinder_sampling(cuda_device):
    from panopticsegforlargescalepointcloud_b200 import transforms as T, _lib
    rng = np.random.default_rng(3)
    pos = rng.uniform(-20, 20, (30000, 3)).astype(np.float32)
    lab = rng.integers(0, 5, 30000)
    ind = tr.cylinder(pos, (2.0, 3.0), 8.0)
    out = T.CylinderSampling(8.0, np.array([2.0, 3.0, 0.0]))({"pos": torch.from_numpy(pos).to(cuda_device), "y": torch.from_numpy(lab).to(cuda_device)})
    want = pos[ind].copy(); want[:, :2] -= np.array([2.0, 3.0], np.float32)
    assert np.array_equal(out["pos"].cpu().numpy(), want) and np.array_equal(out["y"].cpu().numpy(), lab[ind])
    with pytest.raises(_lib.PgsError):
        T.CylinderSampling(8.0, np.array([0.0, 0.0]))({"pos": torch.from_numpy(pos)})
