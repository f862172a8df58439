"""Seeded synthetic cylinder samples in the shape of the reference's two datasets.

There is no dataset on the GPU box, so bench.py / smoke() / the tests feed scenes made here.  The output
follows the reference's input contract:

  pos f32[N,3], coords i32[N,3] = round(pos / grid) with one point per voxel
      (torch_points3d/core/data_transform/grid_transform.py:181-198),
  x f32[N,4] = [x - mean, y - mean, z - mean, z]  (core/data_transform/features.py:391-397),
  y i64[N], instance_labels i64[N], instance_mask bool[N], vote_label f32[N,3], num_instances
      (torch_points3d/datasets/panoptic/utils.py:4-49).

kind="urban":  NPM3D-shape, 9 classes, things {2,3,4,6,7,8} (datasets/panoptic/npm3d.py:18-29,48)
kind="forest": FOR-instance-shape, 2 classes, thing {1}     (datasets/panoptic/treeins.py:21-36)

Pure numpy, CPU; this is data generation, not part of the measured path.
"""
import numpy as np

URBAN_THINGS = (2, 3, 4, 6, 7, 8)
URBAN_STUFF = (0, 1, 5)
FOREST_THINGS = (1,)
FOREST_STUFF = (0,)


def num_classes(kind):
    return 9 if kind == "urban" else 2


def stuff_classes(kind):
    return URBAN_STUFF if kind == "urban" else FOREST_STUFF


class Scene:
    """Plain record; attribute names are the reference Batch keys."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def keys(self):
        return list(self.__dict__.keys())


def _rot(rng, pts):
    a = rng.uniform(0, 2 * np.pi)
    c, s = np.cos(a), np.sin(a)
    out = pts.copy()
    out[:, 0] = c * pts[:, 0] - s * pts[:, 1]
    out[:, 1] = s * pts[:, 0] + c * pts[:, 1]
    return out


def _disc(rng, n, R):
    r = R * np.sqrt(rng.random(n))
    t = rng.uniform(0, 2 * np.pi, n)
    return np.stack([r * np.cos(t), r * np.sin(t)], 1)


def _cyl_surface(rng, n, r, h):
    t = rng.uniform(0, 2 * np.pi, n)
    return np.stack([r * np.cos(t), r * np.sin(t), rng.uniform(0, h, n)], 1)


def _ellipsoid(rng, n, a, b, c, fill=0.25):
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
    rad = np.where(rng.random(n) < fill, rng.random(n) ** (1 / 3), 1.0)[:, None]
    return v * rad * np.array([a, b, c])


def _box_surface(rng, n, lx, ly, lz):
    p = rng.uniform(-0.5, 0.5, (n, 3))
    face = rng.integers(0, 3, n)
    sign = np.where(rng.random(n) < 0.5, -0.5, 0.5)
    p[np.arange(n), face] = sign
    p *= np.array([lx, ly, lz])
    p[:, 2] += lz / 2
    return p


def _n_for(area, grid, density, over=2.5):
    return max(int(over * density * area / (grid * grid)), 12)


class _Builder:
    """Collects labelled surface samples; voxel de-duplication happens once, at the end."""

    def __init__(self, rng, grid, R, shape, density):
        self.rng, self.grid, self.R, self.shape = rng, grid, R, shape
        self.d = min(density, 1.0)          # per-object sampling density (1 = full surface coverage)
        self.mult = max(density, 1.0)       # > 1: more objects instead
        self.pos, self.sem, self.ins = [], [], []
        self.next_instance = 1

    def n_for(self, area, over=2.5):
        return _n_for(area, self.grid, self.d, over)

    def _inside(self, p):
        if self.shape == "disc":
            return p[:, 0] ** 2 + p[:, 1] ** 2 <= self.R * self.R
        return (np.abs(p[:, 0]) <= self.R) & (np.abs(p[:, 1]) <= self.R)

    def add(self, p, sem, thing):
        p = p[self._inside(p)].astype(np.float32)
        if len(p) == 0:
            return
        iid = 0
        if thing:
            iid = self.next_instance
            self.next_instance += 1
        self.pos.append(p)
        self.sem.append(np.full(len(p), sem, np.int64))
        self.ins.append(np.full(len(p), iid, np.int64))

    def place(self, margin=0.0):
        if self.shape == "disc":
            return _disc(self.rng, 1, max(self.R - margin, 0.1))[0]
        return self.rng.uniform(-self.R + margin, self.R - margin, 2)

    def count(self, per_800m2, area, lo=1):
        return max(int(round(per_800m2 * area / 800.0 * self.mult)), lo)


def _tree(b, rng, sem, h_lo, h_hi):
    h = rng.uniform(h_lo, h_hi)
    tr = rng.uniform(0.1, 0.4)
    ca, cc = rng.uniform(1.5, 4.0), rng.uniform(0.25, 0.45) * h
    xy = b.place(1.0)
    trunk = _cyl_surface(rng, b.n_for(2 * np.pi * tr * (h - cc)), tr, h - cc)
    area = 4 * np.pi * ((ca * ca) ** 1.6 * 2 / 3 + (ca * cc) ** 1.6 / 3) ** (1 / 1.6)
    crown = _ellipsoid(rng, b.n_for(area * 1.6), ca, ca, cc, fill=0.4)
    crown[:, 2] += h - cc
    p = np.concatenate([trunk, crown])
    p[:, :2] += xy
    b.add(p, sem, True)


def _voxel_keys(pos, grid):
    c = np.round(pos / np.float32(grid)).astype(np.int64)
    return c, ((c[:, 0] + 32768) << 32) | ((c[:, 1] + 32768) << 16) | (c[:, 2] + 32768)


def _generate(kind, grid, radius, shape, seed, density):
    rng = np.random.default_rng(seed)
    b = _Builder(rng, grid, radius, shape, density)
    area = np.pi * radius ** 2 if shape == "disc" else 4 * radius ** 2
    n_g = b.n_for(area)
    gxy = _disc(rng, n_g, radius) if shape == "disc" else rng.uniform(-radius, radius, (n_g, 2))
    b.add(np.concatenate([gxy, rng.normal(0, 0.05, (n_g, 1))], 1), 0, False)          # ground (stuff 0)
    if kind == "urban":
        small = [(2, 0.10, (4, 8)), (3, 0.10, (0.8, 1.1)), (4, 0.30, (0.9, 1.2))]
        for _ in range(b.count(12, area, 3)):              # poles / bollards / trash cans
            sem, r, (h0, h1) = small[rng.integers(0, 3)]
            h = rng.uniform(h0, h1)
            p = _cyl_surface(rng, b.n_for(2 * np.pi * r * h, 4), r, h)
            p[:, :2] += b.place(0.5)
            b.add(p, sem, True)
        for _ in range(b.count(10, area, 2)):              # pedestrians
            p = _ellipsoid(rng, b.n_for(4.0, 3), 0.25, 0.25, 0.9, fill=0.0)
            p[:, 2] += 0.9
            p[:, :2] += b.place(0.5)
            b.add(p, 6, True)
        for _ in range(b.count(14, area, 2)):              # cars
            p = _rot(rng, _box_surface(rng, b.n_for(30.0), 4.0, 1.8, 1.5))
            p[:, :2] += b.place(2.0)
            b.add(p, 7, True)
        for _ in range(b.count(3, area, 1)):               # barriers (stuff 5)
            L = rng.uniform(5, 15)
            n = b.n_for(L * 1.0)
            p = np.stack([rng.uniform(-L / 2, L / 2, n), rng.normal(0, 0.03, n), rng.uniform(0, 1.0, n)], 1)
            p = _rot(rng, p)
            p[:, :2] += b.place(1.0)
            b.add(p, 5, False)
        for _ in range(b.count(6, area, 3)):               # facades (stuff 1)
            L, H = rng.uniform(10, 30), rng.uniform(6, 15)
            n = b.n_for(L * H)
            p = np.stack([rng.uniform(-L / 2, L / 2, n), rng.normal(0, 0.02, n), rng.uniform(0, H, n)], 1)
            p = _rot(rng, p)
            p[:, :2] += b.place(0.0)
            b.add(p, 1, False)
        for _ in range(b.count(12, area, 3)):              # trees (thing 8)
            _tree(b, rng, 8, 5, 14)
    elif kind == "forest":
        n_u = b.n_for(0.1 * area)
        uxy = _disc(rng, n_u, radius) if shape == "disc" else rng.uniform(-radius, radius, (n_u, 2))
        b.add(np.concatenate([uxy, rng.uniform(0.05, 1.5, (n_u, 1))], 1), 0, False)   # understory
        for _ in range(b.count(160, area, 8)):             # 20-60 trees on an 8 m cylinder
            _tree(b, rng, 1, 10, 35)
    else:
        raise ValueError("kind must be 'urban' or 'forest'")
    pos = np.concatenate(b.pos)
    sem = np.concatenate(b.sem)
    ins = np.concatenate(b.ins)
    # Center (conf/data/panoptic/treeins_rad8.yaml:53) then one point per voxel, first sample wins
    pos = (pos - pos.mean(0, keepdims=True)).astype(np.float32)
    _, key = _voxel_keys(pos, grid)
    _, first = np.unique(key, return_index=True)
    return pos[first], sem[first], ins[first], rng


def make_scene(kind="urban", n_target=200000, grid=0.12, radius=16.0, seed=0, shape="disc", batch_id=0):
    """-> Scene with exactly n_target rows (the sampling density is adapted until the geometry yields enough)."""
    density = 0.5
    for _ in range(8):
        pos, sem, ins, rng = _generate(kind, grid, radius, shape, seed, density)
        if len(pos) >= n_target:
            break
        density *= max(1.15 * n_target / max(len(pos), 1), 1.15)
    n = min(n_target, len(pos))
    sel = rng.permutation(len(pos))[:n]                    # random row order, like a shuffled voxel set
    pos, sem, ins = pos[sel], sem[sel], ins[sel]
    coords = _voxel_keys(pos, grid)[0].astype(np.int32)

    # relabel instances 1..M; 0 = stuff (datasets/panoptic/utils.py:4-49)
    uniq, inv = np.unique(ins, return_inverse=True)
    if uniq[0] != 0:
        inv = inv + 1
    ins = inv.astype(np.int64)
    m = int(ins.max())
    vote = np.zeros((n, 3), np.float32)
    order = np.argsort(ins, kind="stable")
    bounds = np.searchsorted(ins[order], np.arange(m + 2))
    centres = np.zeros((m, 3), np.float32)
    for i in range(1, m + 1):
        rows = order[bounds[i]:bounds[i + 1]]
        c = 0.5 * (pos[rows].min(0) + pos[rows].max(0))     # bbox centre (utils.py:28-32)
        centres[i - 1] = c
        vote[rows] = c - pos[rows]
    x = np.concatenate([pos - pos.mean(0, keepdims=True), pos[:, 2:3]], 1).astype(np.float32)
    return Scene(pos=pos, coords=coords, x=x, y=sem, instance_labels=ins, instance_mask=ins > 0,
                 vote_label=vote, center_label=centres, num_instances=np.array([m], np.int64),
                 batch=np.full(n, batch_id, np.int64), grid_size=np.float32(grid), kind=kind)


def collate(scenes):
    """Batch.from_data_list for the keys the hot path reads (datasets/base_dataset.py:174)."""
    out = {}
    for k in ("pos", "coords", "x", "y", "instance_mask", "vote_label", "center_label", "num_instances"):
        out[k] = np.concatenate([getattr(s, k) for s in scenes])
    out["batch"] = np.concatenate([np.full(len(s.pos), i, np.int64) for i, s in enumerate(scenes)])
    # instance ids stay per-scene (the losses loop over scenes: panoptic_losses.py:203-343)
    out["instance_labels"] = np.concatenate([s.instance_labels for s in scenes])
    out["grid_size"] = scenes[0].grid_size
    out["kind"] = scenes[0].kind
    return Scene(**out)


def synthetic_head_outputs(scene, seed=0, embed_dim=5, label_noise=0.02):
    """Head outputs a trained model would produce, used to drive the clustering stage:
    offset = (centre - pos) + N(0, 0.3 g), embed = mu_inst + N(0, 0.15), semantic = one-hot(y) with label noise."""
    rng = np.random.default_rng(10_000 + seed)
    n = len(scene.pos)
    g = float(scene.grid_size)
    nc = num_classes(scene.kind)
    offset = scene.vote_label + rng.normal(0, 0.3 * g, (n, 3)).astype(np.float32)
    offset[~scene.instance_mask] = rng.normal(0, 0.3 * g, (int((~scene.instance_mask).sum()), 3))
    # instance ids may repeat across scenes of a batch: key the means by (batch, instance)
    gid = scene.batch * (int(scene.instance_labels.max()) + 1) + scene.instance_labels
    uniq, inv = np.unique(gid, return_inverse=True)
    mu = rng.normal(0, 3.0, (len(uniq), embed_dim))
    embed = (mu[inv] + rng.normal(0, 0.15, (n, embed_dim))).astype(np.float32)
    y = scene.y.copy()
    flip = rng.random(n) < label_noise
    y[flip] = rng.integers(0, nc, int(flip.sum()))
    logits = np.full((n, nc), -10.0, np.float32)
    logits[np.arange(n), y] = 0.0
    return offset.astype(np.float32), embed, logits
