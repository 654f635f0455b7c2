"""The C++ drop-in headers (include/MA/*.hpp over the C-ABI): the reference's driver flows compiled
with g++ against this repository's include/MA and run on the GPU, checked against the same calls made
through the Python mirror and against the invariants the reference's drivers print."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_dropin")


def build_driver():
    from mongeampere_b200 import build as b
    b.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "include", "MA", f) for f in os.listdir(os.path.join(ROOT, "include", "MA"))]
    deps.append(os.path.join(ROOT, "include", "ma_b200.h"))
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    libdir = os.path.join(ROOT, "mongeampere_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                           "-L", libdir, "-lma_b200", "-Wl,-rpath," + libdir, "-o", EXE])
    return EXE


def test_headers_compile_and_fail_loudly_without_gpu():
    exe = build_driver()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "50", "8"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


def glibc_points(N):
    """tests/test_opttransport.cpp:19-22,35-42: rand() without srand."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    X = np.empty((N, 2))
    for i in range(N):
        X[i, 0] = 0.999 * (2 * (libc.rand() / (2147483647 + 1.0)) - 1)
        X[i, 1] = 0.999 * (2 * (libc.rand() / (2147483647 + 1.0)) - 1)
    return X


def image(n):
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    x, y = -1 + 2.0 * i / (n - 1), -1 + 2.0 * j / (n - 1)
    v = 200 * np.exp(-((x - 0.3) ** 2 + (y + 0.2) ** 2) / 0.08) + 120 * np.exp(-((x + 0.4) ** 2 + (y - 0.4) ** 2) / 0.02)
    return np.floor(np.minimum(v, 255.0))  # [i, j]


@pytest.mark.gpu
def test_cpp_dropin_matches_c_abi(gpu_ctx):
    exe = build_driver()
    N, n = 2000, 48
    r = subprocess.run([exe, str(N), str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = {}
    for line in r.stdout.splitlines():
        k, *v = line.split()
        out[k] = [float(a) for a in v]
    # the same problem through the Python mirror of the C-ABI
    img = image(n)
    tm = gpu_ctx.set_image(img)  # img[i, j] = image(i, j)
    X = glibc_points(N)
    gpu_ctx.set_points(X)
    assert abs(out["total_mass"][0] - tm) <= 1e-12 * tm
    f0, g0, H0 = gpu_ctx.kantorovich(np.zeros(N))
    assert abs(out["f0"][0] - f0) <= 1e-11 * abs(f0)
    assert abs(out["sum_g0"][0] - tm) <= 1e-11 * tm
    assert int(out["nnz0"][0]) == H0.nnz
    assert out["max_rowsum0"][0] <= 1e-9 * abs(H0.diagonal()).max()
    assert out["laplace_residual"][0] <= 1e-7 and out["laplace_last"][0] == 0.0
    nu = np.full(N, tm / N)
    w, st, rc = gpu_ctx.ot_solve(nu, eps_g=1e-9, maxiter=100, verbose=False)
    assert rc == 0
    assert int(out["niter"][0]) == st["niter"] and int(out["neval"][0]) == st["neval"]
    assert out["final_norm"][0] < 1e-9
    assert abs(out["w_first"][0] - w[0]) <= 1e-9 * np.abs(w).max() and abs(out["w_last"][0] - w[-1]) <= 1e-9 * np.abs(w).max()
    f1, g1, _ = gpu_ctx.kantorovich(w)
    assert abs(out["f_final"][0] - f1) <= 1e-10 * abs(f1)
    m, c = gpu_ctx.lloyd(w)
    assert abs(out["lloyd_mass_sum"][0] - tm) <= 1e-10 * tm
    assert np.allclose(out["lloyd_c0"], c[0], rtol=0, atol=1e-10)
    # tests/test_voronoi_tri.cpp:67: the pieces tile the domain; cell 0's pieces have the cell's area
    assert abs(out["area_sum"][0] - 4.0) <= 1e-11
    assert out["pieces"][0] >= N
    # voronoi_triangulation_intersection_raw (vti.hpp:219-313): the symbolic polygons rebuilt from their EDGE_T / EDGE_DT
    # tags are the same pieces
    assert out["raw_pieces"][0] == out["pieces"][0] and out["raw_edge_t"][0] > 0 and out["raw_edge_dt"][0] > 0
    assert abs(out["raw_area_sum"][0] - 4.0) <= 1e-9
    # draw_laguerre_diagram with colour 1: the image integrates (mean vertex density per piece), close to the total mass
    assert abs(out["raster_sum"][0] - tm) <= 0.05 * tm
    # tests/test_power.cpp:51: the cells clipped to a convex polygon tile it
    assert abs(out["cells_area_sum"][0] - out["pentagon_area"][0]) <= 1e-12
    # tests/test_voronoi_ad.cpp:60: the same with a non-convex polygon (a cross)
    assert abs(out["cross_area"][0] - (4 * 0.3 * 0.9 + 4 * 0.3 * 0.6)) <= 1e-14  # 2a*2b + 2*(b-a)*2a
    assert abs(out["cross_cells_area_sum"][0] - out["cross_area"][0]) <= 1e-12
    # the image handed over as an explicit triangulation: the same diagonals give the same numbers (ma_set_mesh recognises
    # the grid), alternating diagonals another PL density whose cells carry its own total mass
    assert abs(out["f0_explicit"][0] - out["f0"][0]) <= 1e-12 * abs(out["f0"][0])
    assert out["nnz0_explicit"][0] == out["nnz0"][0]
    assert abs(out["sum_g0_explicit"][0] - out["tm_explicit"][0]) <= 1e-11 * tm
    assert abs(out["tm_explicit"][0] - tm) <= 1e-12 * tm
    assert abs(out["sum_g0_alternating"][0] - out["tm_alternating"][0]) <= 1e-11 * tm
    assert out["nnz0_alternating"][0] == out["nnz0"][0]  # the Laguerre adjacency does not depend on the density


def test_header_layer_host_parts():
    """include/MA/functions.hpp + lite.hpp without any GPU call (tests/cpp/test_lite_host.cpp)."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_lite_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_lite_host.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
