"""Synthetic inputs of the five BASELINE.json configs, numpy only.

Meshes follow the reference's conventions:
  * unit square of tests/test_triangulation.cpp:12-21 (4 vertices, faces {0,1,2},{0,2,3});
  * image grids of include/MA/functions.hpp:93-102: an n x m image gives vertex (i, j) at
    (-1 + 2 i/(n-1), -1 + 2 j/(m-1)), stored at index i*m + j, with density
    image(i, m-j-1)/255 + 1e-3.  The reference lets CGAL pick the diagonal of every grid square
    (not reproducible); here every square (i, j) is split along
    (i, j)-(i+1, j+1) into faces 2*(i*(m-1)+j) = {(i,j),(i+1,j),(i+1,j+1)} and
    2*(i*(m-1)+j)+1 = {(i,j),(i+1,j+1),(i,j+1)}, both counter-clockwise.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
def unit_square_mesh():
    vx = np.array([0.0, 1.0, 1.0, 0.0])
    vy = np.array([0.0, 0.0, 1.0, 1.0])
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return vx, vy, tri


def grid_vertices(n: int, m: int, x0=-1.0, y0=-1.0, x1=1.0, y1=1.0):
    dx = (x1 - x0) / float(n - 1)
    dy = (y1 - y0) / float(m - 1)
    i = np.arange(n, dtype=np.float64)
    j = np.arange(m, dtype=np.float64)
    vx = np.repeat(x0 + i * dx, m)
    vy = np.tile(y0 + j * dy, n)
    return vx, vy


def grid_triangles(n: int, m: int):
    i, j = np.meshgrid(np.arange(n - 1), np.arange(m - 1), indexing="ij")
    i = i.reshape(-1)
    j = j.reshape(-1)
    v00 = i * m + j
    v10 = (i + 1) * m + j
    v11 = (i + 1) * m + j + 1
    v01 = i * m + j + 1
    tri = np.empty((2 * len(i), 3), np.int32)
    tri[0::2] = np.stack([v00, v10, v11], 1)
    tri[1::2] = np.stack([v00, v11, v01], 1)
    return tri


def grid_triangles_diag(n: int, m: int, diag):
    """The same grid with a per-square choice of diagonal: diag[i, j] = 0 splits square (i, j) along (i,j)-(i+1,j+1) as
    grid_triangles does, 1 along (i+1,j)-(i,j+1) — what a Delaunay triangulation of the (co-circular) pixel grid may pick,
    SURVEY App. B T1.  Faces 2q, 2q+1 of square q = i (m-1) + j, CCW."""
    diag = np.asarray(diag).reshape(n - 1, m - 1).astype(bool).reshape(-1)
    tri = grid_triangles(n, m)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(m - 1), indexing="ij")
    i = i.reshape(-1); j = j.reshape(-1)
    v00, v10, v11, v01 = i * m + j, (i + 1) * m + j, (i + 1) * m + j + 1, i * m + j + 1
    alt0 = np.stack([v00, v10, v01], 1).astype(np.int32)
    alt1 = np.stack([v10, v11, v01], 1).astype(np.int32)
    tri[0::2][diag] = alt0[diag]
    tri[1::2][diag] = alt1[diag]
    return tri


def image_vertex_density(image: np.ndarray):
    """image[i, j] = CImg image(i, j) (i = column / x index, j = row index) -> rho at vertex index
    i*m + j, = image(i, m-j-1)/255 + 1e-3 (functions.hpp:102)."""
    n, m = image.shape
    return (image[:, ::-1] / 255.0 + 1e-3).reshape(-1).astype(np.float64)


def pl_coefficients(vx, vy, rho, tri):
    """Per-face (a, b, c) with rho_f(x, y) = a x + b y + c interpolating the vertex values
    (what MA::Linear_function represents, functions.hpp:55-80)."""
    tri = np.asarray(tri).reshape(-1, 3)
    ax, ay, fa = vx[tri[:, 0]], vy[tri[:, 0]], rho[tri[:, 0]]
    bx, by, fb = vx[tri[:, 1]], vy[tri[:, 1]], rho[tri[:, 1]]
    cx, cy, fc = vx[tri[:, 2]], vy[tri[:, 2]], rho[tri[:, 2]]
    det = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay)
    a = ((fb - fa) * (cy - ay) - (fc - fa) * (by - ay)) / det
    b = ((bx - ax) * (fc - fa) - (cx - ax) * (fb - fa)) / det
    c = fa - a * ax - b * ay
    return np.stack([a, b, c], 1)


def total_mass(vx, vy, tri, abc):
    """sum_f area_f * rho_f(centroid_f) (functions.hpp:117)."""
    tri = np.asarray(tri).reshape(-1, 3)
    ax, ay = vx[tri[:, 0]], vy[tri[:, 0]]
    bx, by = vx[tri[:, 1]], vy[tri[:, 1]]
    cx, cy = vx[tri[:, 2]], vy[tri[:, 2]]
    area = ((bx - ax) * (cy - ay) - (cx - ax) * (by - ay)) / 2
    gx, gy = (ax + bx + cx) / 3, (ay + by + cy) / 3
    return float(np.sum(area * (abc[:, 0] * gx + abc[:, 1] * gy + abc[:, 2])))


# ------------------------------------------------------------------------------------------------
# densities
# ------------------------------------------------------------------------------------------------
C2_MIXTURE = ((1.0, (0.3, -0.2), 0.20), (0.5, (-0.4, 0.4), 0.10), (0.8, (-0.5, -0.5), 0.15),
              (0.6, (0.5, 0.6), 0.08))


def gaussian_mixture_density(vx, vy, mixture=C2_MIXTURE, floor=1e-3):
    rho = np.full_like(vx, floor)
    for a, (cx, cy), s in mixture:
        rho += a * np.exp(-((vx - cx) ** 2 + (vy - cy) ** 2) / (2 * s * s))
    return rho


def synthetic_image(n: int, m: int, seed: int = 3, blobs: int = 64):
    """8-bit 'photograph': sum of random Gaussians + smooth noise, quantised to 0..255 (config c3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.linspace(-1, 1, n)[:, None]
    y = np.linspace(-1, 1, m)[None, :]
    img = np.zeros((n, m))
    for _ in range(blobs):
        cx, cy = rng.uniform(-1, 1, 2)
        s = rng.uniform(0.03, 0.25)
        a = rng.uniform(0.2, 1.0)
        img += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))
    for k in range(1, 5):  # smooth noise
        ph = rng.uniform(0, 2 * np.pi, 2)
        img += 0.05 * np.sin(3 * k * np.pi * x + ph[0]) * np.cos(2 * k * np.pi * y + ph[1])
    img -= img.min()
    img /= img.max()
    return np.round(img * 255.0)


# ------------------------------------------------------------------------------------------------
# Diracs
# ------------------------------------------------------------------------------------------------
def uniform_points(N: int, seed: int, lo=(0.0, 0.0), hi=(1.0, 1.0)):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((N, 2))
    X[:, 0] = lo[0] + (hi[0] - lo[0]) * X[:, 0]
    X[:, 1] = lo[1] + (hi[1] - lo[1]) * X[:, 1]
    return X


class GlibcRand:
    """glibc rand() (TYPE_3 additive feedback generator), so that config 1 can use the literal point
    set of tests/test_opttransport.cpp:19-22,35-42 (no srand -> seed 1)."""

    def __init__(self, seed: int = 1):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            if w < 0:
                w += 2147483647
            r[i] = w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = r
        for _ in range(310):
            self._step()

    def _step(self):
        v = (self.r[-31] + self.r[-3]) & 0xFFFFFFFF
        self.r.append(v)
        self.r.pop(0)
        return v

    def rand(self) -> int:
        return self._step() >> 1


def glibc_rr_points(N: int, scale: float = 1.0):
    """X(i,:) = (rr(), rr()) with rr() = 2*rand()/(RAND_MAX+1) - 1 (tests/test_opttransport.cpp:19-22)."""
    g = GlibcRand(1)
    X = np.empty((N, 2))
    for i in range(N):
        X[i, 0] = (2 * (g.rand() / 2147483648.0) - 1) * scale
        X[i, 1] = (2 * (g.rand() / 2147483648.0) - 1) * scale
    return X


def jittered_grid_points(n: int, seed: int):
    """bench_opttransport.cpp:45-60 style: n x n grid on [-1,1]^2 jittered by rr()/(2n)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = 2.0 / (n - 1)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    X = np.empty((n * n, 2))
    X[:, 0] = -1.0 + i.reshape(-1) * d + (2 * rng.random(n * n) - 1) / (2.0 * n)
    X[:, 1] = -1.0 + j.reshape(-1) * d + (2 * rng.random(n * n) - 1) / (2.0 * n)
    return X


# ------------------------------------------------------------------------------------------------
# the five configs
# ------------------------------------------------------------------------------------------------
def config(name: str, scale: float = 1.0):
    """-> dict(kind, mesh description, X, nu).  `scale` shrinks N (and the grid) for tests."""
    if name == "c1":
        N = max(4, int(10_000 * scale))
        vx, vy, tri = unit_square_mesh()
        rho = np.ones(4)
        X = uniform_points(N, 1)
        return dict(kind="mesh", vx=vx, vy=vy, tri=tri, rho=rho, X=X)
    if name == "c1r":  # literal test_opttransport / test_zeldovich style
        N = max(4, int(10_000 * scale))
        vx, vy = grid_vertices(2, 2)
        rho = image_vertex_density(np.full((2, 2), 255.0))
        return dict(kind="grid", n=2, m=2, vx=vx, vy=vy, tri=grid_triangles(2, 2), rho=rho,
                    X=glibc_rr_points(N))
    if name in ("c2", "c4"):
        n = 512 if name == "c2" else 1024
        N = 100_000 if name == "c2" else 250_000
        n = max(4, int(round(n * np.sqrt(scale))))
        N = max(16, int(N * scale))
        vx, vy = grid_vertices(n, n)
        rho = gaussian_mixture_density(vx, vy)
        X = uniform_points(N, 2 if name == "c2" else 6, (-1, -1), (1, 1))
        return dict(kind="grid", n=n, m=n, vx=vx, vy=vy, tri=grid_triangles(n, n), rho=rho, X=X)
    if name == "c3":
        n = max(4, int(round(2048 * np.sqrt(scale))))
        N = max(16, int(1_000_000 * scale))
        vx, vy = grid_vertices(n, n)
        rho = image_vertex_density(synthetic_image(n, n, 3))
        X = uniform_points(N, 4, (-1, -1), (1, 1))
        return dict(kind="grid", n=n, m=n, vx=vx, vy=vy, tri=grid_triangles(n, n), rho=rho, X=X)
    if name == "c5":
        N = max(16, int(4_000_000 * scale))
        vx, vy = grid_vertices(2, 2)
        rho = image_vertex_density(np.full((2, 2), 255.0))
        X = uniform_points(N, 7, (-1 / 1.1, -1 / 1.1), (1 / 1.1, 1 / 1.1))
        return dict(kind="grid", n=2, m=2, vx=vx, vy=vy, tri=grid_triangles(2, 2), rho=rho, X=X)
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------
# image files / irregular meshes
# ------------------------------------------------------------------------------------------------
def read_pgm(path):
    """PGM (P2 ascii / P5 binary, 8 or 16 bit) -> image[i, j] = pixel (column i, row j from the top), the CImg(i, j)
    indexing that image_vertex_density / ma_set_image expect (the reference's drivers load their density with
    cimg_library::CImg<double>(path), tests/test_opttransport.cpp:45-49; CImg itself is not available here)."""
    import re
    with open(path, "rb") as fh:
        data = fh.read()
    m = re.match(rb"(P[25])\s+(?:#[^\n]*\n\s*)*(\d+)\s+(?:#[^\n]*\n\s*)*(\d+)\s+(?:#[^\n]*\n\s*)*(\d+)\s", data)
    if not m:
        raise ValueError(f"{path}: not a PGM file")
    w, h, maxval = int(m.group(2)), int(m.group(3)), int(m.group(4))
    if m.group(1) == b"P2":
        px = np.array(data[m.end():].split(), dtype=np.float64)
    else:
        dt = np.dtype(">u2") if maxval > 255 else np.uint8
        px = np.frombuffer(data, dtype=dt, count=w * h, offset=m.end()).astype(np.float64)
    if px.size != w * h:
        raise ValueError(f"{path}: truncated PGM file")
    return np.ascontiguousarray(px.reshape(h, w).T)


def l_shaped_mesh(nv: int = 400, seed: int = 0):
    """An irregular triangulation of the NON-convex domain [-1,1]^2 minus (0,1]x(0,1] (Delaunay of random interior
    points + points on the boundary, triangles of the removed quadrant dropped) with a random PL density:
    (vx, vy, tri CCW, rho) in the format of ma_set_mesh_pl / Triangulation_incremental_builder_2."""
    from scipy.spatial import Delaunay
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.linspace(-1, 1, 17)
    s = np.linspace(0, 1, 9)
    bnd = np.r_[np.c_[t, -np.ones_like(t)], np.c_[-np.ones_like(t), t], np.c_[t[t <= 0], np.ones((t <= 0).sum())],
                np.c_[np.ones((t <= 0).sum()), t[t <= 0]], np.c_[s, np.zeros_like(s)], np.c_[np.zeros_like(s), s]]
    P = rng.uniform(-1, 1, (3 * nv, 2))
    P = P[~((P[:, 0] > 0.02) & (P[:, 1] > 0.02))]
    P = P[~((P[:, 0] > -0.02) & (P[:, 1] > -0.02) & ((P[:, 0] < 0.02) | (P[:, 1] < 0.02)))][:nv]
    pts = np.unique(np.r_[bnd, P].round(12), axis=0)
    d = Delaunay(pts)
    tri = d.simplices
    c = pts[tri].mean(axis=1)
    tri = tri[~((c[:, 0] > 0) & (c[:, 1] > 0))]
    a, b, cc = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (cc[:, 1] - a[:, 1]) - (cc[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tri = tri[np.abs(area2) > 1e-12]
    area2 = area2[np.abs(area2) > 1e-12]
    tri = np.where(area2[:, None] > 0, tri, tri[:, ::-1]).astype(np.int32)
    rho = 0.3 + rng.random(len(pts)) + np.exp(-((pts[:, 0] + 0.4) ** 2 + (pts[:, 1] + 0.3) ** 2) / 0.08)
    return np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1]), np.ascontiguousarray(tri), rho
