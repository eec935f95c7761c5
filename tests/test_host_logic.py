"""CPU-only checks: the C-ABI library loads and exports every symbol include/fvb200.h declares,
refuses to run without a GPU (no CPU fallback), and the host-side helpers (grid builder, step
controller, slab planning) agree with the oracle / the reference's own tests."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(fv):
    hdr = open(os.path.join(ROOT, "include", "fvb200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char \*)\s*(fvb_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 30
    assert declared == set(fv._lib.SYMBOLS)
    L = ctypes.CDLL(fv.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.fvb_version() >= 100


def test_device_arena_bookkeeping(tmp_path):
    """csrc/arena.h (host-only bookkeeping of the handle's device memory) under ASan/UBSan with malloc-backed
    chunks: steady state after one pass of a repeated request sequence (same addresses, no growth), no
    overlaps, exact accounting, full coalescing, the out-of-memory / trim path."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "test_arena")
    src = os.path.join(ROOT, "tests", "cpp", "test_arena.cpp")
    subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "arena ok" in r.stdout, r.stdout + r.stderr


def test_host_util_helpers(tmp_path):
    """csrc/host_util.h (sorted-unique of the halo references with a bitmap pass, the multi-rank Dirichlet table)
    against the straightforward formulations, under ASan/UBSan."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "test_host_util")
    subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_host_util.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_util ok" in r.stdout, r.stdout + r.stderr


def _dia_tma_model_check(L, offs, unit, nf):
    """Replay, in numpy, what the producer lane of k_spmv_dia_tma copies for every interior tile and what the
    consumer threads read back (csrc/dia_tma.cuh), on the layout the library's own dia_tma_layout() produced.
    Every cp.async.bulk must be 16-byte aligned on both sides, stay inside its array, and the stage must be
    filled exactly once; every consumer read must return the element the row sum needs."""
    T, M, near = L["T"], L["margin"], L["near_max"]
    X = np.arange(nf, dtype=np.int64) * 16 + 1
    DG = np.arange(nf, dtype=np.int64) * 16 + 2
    U = [np.arange(nf + o, dtype=np.int64) * 16 + 3 + k for k, o in enumerate(offs)]
    even_up = lambda o: (o + 1) & ~1
    n_interior = 0
    for tile in range((nf + T - 1) // T):
        r0 = tile * T
        if not (r0 >= L["reach"] and r0 + T + L["reach"] <= nf):
            continue
        n_interior += 1
        sm = np.full(L["stage_doubles"], -1, dtype=np.int64)
        copied = 0

        def cp(dst, arr, start, n):
            nonlocal copied
            assert start % 2 == 0 and n % 2 == 0 and dst % 2 == 0      # 16-byte alignment, size multiple of 16
            assert 0 <= start and start + n <= arr.size and dst + n <= sm.size
            assert np.all(sm[dst:dst + n] == -1)                       # no two copies overlap in the stage
            sm[dst:dst + n] = arr[start:start + n]
            copied += n
        cp(L["xc"], X, r0 - M, T + 2 * M)
        if not unit:
            cp(L["dg"], DG, r0, T)
        for k, o in enumerate(offs):
            if o <= near:
                cp(L["un"][k], U[k], r0, T + even_up(o))
            else:
                sh = o & 1
                cp(L["xl"][k], X, r0 - o - sh, T + 2)
                cp(L["xu"][k], X, r0 + o - sh, T + 2)
                cp(L["ul"][k], U[k], r0, T)
                cp(L["uu"][k], U[k], r0 + o - sh, T + 2)
        assert copied * 8 == L["tx_bytes"] and np.all(sm != -1)
        i = np.arange(T)
        r = r0 + i
        assert np.array_equal(sm[L["xc"] + M + i], X[r])
        if not unit:
            assert np.array_equal(sm[L["dg"] + i], DG[r])
        for k, o in enumerate(offs):
            if o <= near:
                lo, xlo = sm[L["un"][k] + i], sm[L["xc"] + M + i - o]
                up, xup = sm[L["un"][k] + i + o], sm[L["xc"] + M + i + o]
            else:
                sh = o & 1
                lo, xlo = sm[L["ul"][k] + i], sm[L["xl"][k] + i + sh]
                up, xup = sm[L["uu"][k] + i + sh], sm[L["xu"][k] + i + sh]
            assert np.array_equal(lo, U[k][r]) and np.array_equal(xlo, X[r - o])         # A[r, r-o] * x[r-o]
            assert np.array_equal(up, U[k][o + r]) and np.array_equal(xup, X[r + o])     # A[r, r+o] * x[r+o]
    return n_interior


def test_dia_tma_layout_addressing(tmp_path):
    """Shared-memory layout of the TMA diagonal SpMV (host function dia_tma_layout in csrc/dia_tma.cuh, dumped by
    a host-only nvcc harness) against a numpy replay of the kernel's copies and reads: near/far, odd/even
    offsets, 1 to 4 diagonals, with and without the stored main diagonal."""
    import json
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "dump_layout")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "dump_dia_tma_layout.cu")], check=True)
    cases = [[1, 512, 262144], [1, 256, 65536], [1, 2, 200], [1, 7, 231], [1, 9, 45], [1], [1, 64, 4096],
             [1, 1024, 1048576], [3, 513, 9999], [1, 2, 4], [511, 512, 513, 700]]
    for offs in cases:
        for unit in (0, 1):
            out = subprocess.run([exe, str(unit)] + [str(o) for o in offs], capture_output=True, text=True, check=True).stdout
            L = json.loads(out)
            assert L["stages"] >= 2 and L["smem_bytes"] <= L["smem_max"] and L["bar_off"] == L["stages"] * L["stage_doubles"] * 8
            assert L["tx_bytes"] == L["stage_doubles"] * 8 and L["bar_off"] % 16 == 0
            nf = min(max(offs) * 3 + 5 * L["T"] + 77, max(offs) * 2 + 40 * L["T"] + 77)
            assert _dia_tma_model_check(L, offs, bool(unit), nf) >= 1
    # the 512^3 steady solve: three stages of the unit-diagonal copy must fit (what the bench runs)
    L = json.loads(subprocess.run([exe, "1", "1", "512", "262144"], capture_output=True, text=True, check=True).stdout)
    assert L["stages"] == 3 and L["margin"] == 512


def _build_c_demo(tmp_path, fv):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(fv.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, fv.LIB_PATH, f"-Wl,-rpath,{libdir}", "-lm"],
                   check=True)
    return exe


def test_header_is_plain_c_and_fails_loudly_without_gpu(fv, tmp_path):
    """include/fvb200.h compiles as pedantic C99 and the library links into a C program (the boundary has no
    C++ or torch types); without a CUDA device that program stops at fvb_create with the no-fallback error."""
    import subprocess
    import torch
    exe = _build_c_demo(tmp_path, fv)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (see test_c_abi_demo_on_gpu)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_abi_demo_on_gpu(fv, tmp_path):
    """examples/c_abi_demo.c: the reference's 4-node smoke test (test/runtests.jl:4-16) and a homogeneous box
    with its exact linear solution, driven from plain C through the C ABI only."""
    import subprocess
    exe = _build_c_demo(tmp_path, fv)
    r = subprocess.run([exe, "24"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_abi_demo ok" in r.stdout, r.stdout + r.stderr


def test_no_cpu_fallback(fv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fv.FVBError, match="no CPU fallback") as e:
        fv.System(0)
    assert e.value.status == 2
    with pytest.raises(fv.FVBError):
        fv.solvediffusion([(1, 2)], [1.0], [1.0], [0.0, 0.0], [1], [0.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "finitevolume.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fv_oracle" not in txt and "oracle/" not in txt, f


@pytest.mark.parametrize("mins,maxs,ns", [([0, 0, 0], [3, 2, 1], [4, 3, 2]),
                                          ([-50, -50, 0], [50, 50, 10], [20, 10, 2]),
                                          ([0, 0, 0], [6, 5, 4], [7, 6, 5])])
def test_regulargrid_matches_oracle(fv, orc, mins, maxs, ns):
    """src/grid.jl:56-110: identical ordering and values as the serial triple loop."""
    c, nb, aol, vol = fv.regulargrid(mins, maxs, ns)
    co, nbo, aolo, volo = orc.regulargrid(mins, maxs, ns)
    assert np.array_equal(nb, nbo) and np.array_equal(aol, aolo) and np.array_equal(vol, volo)
    assert np.allclose(c, co, rtol=0, atol=1e-12)
    assert fv.grid_sizes(ns) == (int(np.prod(ns)), nb.shape[0])


def test_regulargrid_rejects_non_3d(fv):
    with pytest.raises(ValueError, match="only 3 dimensions supported"):  # src/grid.jl:59
        fv.regulargrid([0, 0], [1, 1], [2, 2])


def test_regulargrid_slabs_cover_global_list(fv):
    """Each rank's slab list = the faces touching its planes, in global order."""
    import importlib
    dist = importlib.import_module("fvb200.distributed")
    ns = [9, 4, 3]
    _, nb, aol, vol = fv.regulargrid([0, 0, 0], [8, 3, 2], ns)
    for nranks in (1, 2, 4):
        planes = dist.slab_planes(ns[0], nranks)
        assert planes[0][0] == 1 and planes[-1][1] == ns[0]
        vols = []
        for pl in planes:
            lo, hi = dist.node_range_of_planes(pl, ns[1], ns[2])
            _, nbs, aols, vs = fv.regulargrid([0, 0, 0], [8, 3, 2], ns, planes=pl)
            touch = ((nb[:, 0] >= lo) & (nb[:, 0] <= hi)) | ((nb[:, 1] >= lo) & (nb[:, 1] <= hi))
            assert np.array_equal(nbs, nb[touch]) and np.array_equal(aols, aol[touch])
            vols.append(vs)
        assert np.array_equal(np.concatenate(vols), vol)


def test_slab_planes_balance_free_planes(fv):
    import importlib
    dist = importlib.import_module("fvb200.distributed")
    pl = dist.slab_planes(512, 8)
    free = [hi - lo + 1 - (1 if r in (0, 7) else 0) for r, (lo, hi) in enumerate(pl)]
    assert max(free) - min(free) <= 1 and sum(hi - lo + 1 for lo, hi in pl) == 512
    assert dist.slab_planes(4, 4) == [(1, 1), (2, 2), (3, 3), (4, 4)]
    with pytest.raises(ValueError):
        dist.slab_planes(3, 4)


def test_nodehycos2neighborhycos(fv, orc):
    """src/grid.jl:14-33: geometric mean / arithmetic mean of logs, (n3,n2,n1) layout."""
    ns = [4, 3, 5]
    _, nb, _, _ = fv.regulargrid([0, 0, 0], [3, 2, 4], ns)
    k = np.random.default_rng(0).random((ns[2], ns[1], ns[0])) + 0.1
    for lg in (False, True):
        assert np.array_equal(fv.nodehycos2neighborhycos(nb, k, lg), orc.nodehycos2neighborhycos(nb, k, lg))
    f1 = fv.nodehycos2neighborhycos(nb, k)[0]
    assert math.isclose(f1, math.sqrt(k[0, 0, 0] * k[0, 0, 1]))  # face 1 = node 1 => node 1+n2*n3


def test_ode_controller_with_linearsolver_hook(fv):
    """test/ode.jl:8-40 -- the generic integrator with a caller-supplied linear solver."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    v = np.array([1.0, 2.0, 3.0])
    ys, ts = fv.backwardeulerintegrate_generic(np.ones(3), sp.diags(v).tocsr(), np.zeros(3), 1e-4, 0.0, 0.25, atol=1e-8,
                                               linearsolver=lambda A, b, x0: spla.spsolve(A.tocsc(), b))
    assert ts[-1] == 0.25
    for y, t in zip(ys, ts):
        assert np.allclose(y, np.exp(-v * t), atol=1e-4)
    A = -np.array([[0.5, -1.0], [1.0, -1.0]])

    def exact(t, c1=1, c2=2):
        s7 = math.sqrt(7)
        a, b_ = np.array([1, 0.75]), np.array([0, -s7 / 4])
        return (c1 * math.exp(-t / 4) * (a * math.cos(s7 * t / 4) - b_ * math.sin(s7 * t / 4))
                + c2 * math.exp(-t / 4) * (a * math.sin(s7 * t / 4) + b_ * math.cos(s7 * t / 4)))
    ys, ts = fv.backwardeulerintegrate_generic(exact(0), A, np.zeros(2), 1e-4, 0.0, 0.5, atol=1e-8,
                                               linearsolver=lambda M, b, x0: np.linalg.solve(M, b))
    for y, t in zip(ys, ts):
        assert np.allclose(y, exact(t), atol=1e-4)
    with pytest.raises(RuntimeError, match="no CPU solver"):
        fv.backwardeulerintegrate_generic(np.ones(3), sp.diags(v).tocsr(), np.zeros(3), 1e-4, 0.0, 2.0)


def test_odeadjoint_closed_form(fv):
    """test/odeadjoint.jl:5-41 -- dx/dt = b x, x(0) = a, G = int x dt: forward solution, adjoint solution
    lambda(t) = (1 - exp(b (T - t))) / b and the gradient [ (e^{bT}-1)/b, -a (e^{bT}-1)/b^2 + a T e^{bT}/b ],
    all within the reference's rtol 1e-4, through the generic integrator / adjoint / gradient entry points with the
    reference's own `linearsolver(A, b, x0) = A \\ b`."""
    a, b, T = 1.0, 2.0, 1.0
    A = np.array([[b]])
    solver = lambda M, r, x0: np.linalg.solve(M, r)
    xs, ts_x = fv.backwardeulerintegrate_generic(np.array([a]), -A, lambda t: np.zeros(1), 1e-5, 0.0, T, linearsolver=solver,
                                                 atol=1e-8)
    assert np.allclose(np.array(xs)[:, 0], a * np.exp(b * np.array(ts_x)), rtol=1e-4)
    lambdas, ts_l = fv.adjointintegrate_generic(-A, lambda t: -np.ones(1), (0.0, T), dt0=1e-5, linearsolver=solver, atol=1e-8)
    assert ts_l[0] == 0.0 and ts_l[-1] == T
    assert np.allclose(np.array(lambdas)[:, 0], (1 - np.exp(b * (T - np.array(ts_l)))) / b, rtol=1e-4, atol=1e-9)
    xc = fv.getcontinuoussolution(xs, ts_x)
    lambdac = fv.getcontinuoussolution(lambdas, ts_l)
    dx0dp = np.array([[-1.0], [0.0]])                                  # odeadjoint.jl:21-26
    dfdp = lambda t: np.array([[0.0], [-xc(t)[0]]])                     # odeadjoint.jl:27-31
    grad = fv.gradientintegrate_generic(lambdac, dx0dp, lambda t: np.zeros(2), dfdp, (0.0, T), [ts_x, ts_l])
    exact = np.array([(math.exp(b * T) - 1) / b, -a / b ** 2 * (math.exp(b * T) - 1) + a / b * math.exp(b * T) * T])
    assert np.allclose(grad, exact, rtol=1e-4)


def test_controller_matches_oracle_trajectory(fv, orc):
    """The step-doubling controller (src/transient.jl:78-154) restated twice -- product host code
    vs oracle -- must take the same accepted steps when both use exact solves."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(1)
    n = 12
    M = sp.random(n, n, 0.3, random_state=2)
    A = (M @ M.T + sp.diags(rng.random(n) + 0.5)).tocsr()
    b = rng.random(n)
    u0 = rng.random(n)
    ys, ts = fv.backwardeulerintegrate_generic(u0, A, b, 0.5, 0.0, 7.0, atol=1e-3,
                                               linearsolver=lambda S, r, x0: spla.spsolve(S.tocsc(), r))
    uso, tso = orc.backwardeulerintegrate_core(u0, A, b, 0.5, 0.0, 7.0, atol=1e-3,
                                               linearsolver=lambda A_, dt, r, x0: spla.spsolve(
                                                   (A_ + sp.identity(n) / dt).tocsc(), r))
    assert ts == tso
    assert np.allclose(np.array(ys), np.array(uso), rtol=1e-10)


def test_halo_plan_pure(fv):
    import importlib
    dist = importlib.import_module("fvb200.distributed")
    ranges = [(1, 10), (11, 10), (21, 5)]
    halos = [np.array([11, 12]), np.array([9, 10, 21]), np.array([20])]
    p0 = dist.halo_plan_from_ranges(0, ranges, halos)
    p1 = dist.halo_plan_from_ranges(1, ranges, halos)
    p2 = dist.halo_plan_from_ranges(2, ranges, halos)
    assert p0[0] == [1] and p0[1] == [2] and list(p0[2]) == [8, 9] and p0[3] == [2]
    assert p1[0] == [0, 2] and p1[1] == [2, 1] and list(p1[2]) == [0, 1, 9] and p1[3] == [2, 1]
    assert p2[0] == [1] and p2[1] == [1] and list(p2[2]) == [0] and p2[3] == [1]
    with pytest.raises(ValueError):
        dist.halo_plan_from_ranges(0, ranges, [np.array([5]), halos[1], halos[2]])


REF_FRACTURES = "/root/reference/examples/fractures/fourfractures"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_FRACTURES, "mesh.jld")),
                    reason="the reference checkout is only present in the build container")
def test_jld_reader_against_fixture(fv, fourfractures):
    """fv.jld.load == JLD.load of examples/fractures/ex.jl:9: every variable of the reference's mesh.jld (root
    group in dense link storage: fractal heap) and of pflotran_solution.jld (symbol-table group) must equal the
    committed fixture, which tests/golden/make_fixtures.py extracted independently by absolute file offsets."""
    path = os.path.join(REF_FRACTURES, "mesh.jld")
    want = ["xs", "ys", "zs", "neighbors", "areasoverlengths", "fractureindices", "dirichletnodes", "dirichletheads",
            "conductivities"]
    assert fv.jld.names(path) == sorted(want)
    got = fv.jld.load(path, *want)  # same call shape as the reference's JLD.load(path, names...)
    for name, arr in zip(want, got):
        assert arr.dtype == fourfractures[name].dtype and np.array_equal(arr, fourfractures[name]), name
    assert got[3].shape == (6314, 2) and got[3].flags["C_CONTIGUOUS"]  # ready for the C ABI's 2F interleaved int64
    h = fv.jld.load(os.path.join(REF_FRACTURES, "pflotran_solution.jld"), "h")
    assert np.array_equal(h, fourfractures["pflotran_h"])
    with pytest.raises(KeyError):
        fv.jld.load(path, "nope")


def test_jld_reader_rejects_foreign_files(fv, tmp_path):
    p = tmp_path / "x.jld"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(fv.jld.JLDFormatError):
        fv.jld.load(str(p))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_FRACTURES, "mesh.jld")),
                    reason="the reference checkout is only present in the build container")
def test_fractures_example_inputs(fv, fourfractures):
    """examples/fractures.py (= examples/fractures/ex.jl:6-17): the host side up to the solve."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fractures_example", os.path.join(ROOT, "examples", "fractures.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    p = m.load_problem(fv, REF_FRACTURES)
    assert np.array_equal(p["neighbors"], fourfractures["neighbors"]) and not p["sources"].any()
    f2f = m.fracture_of_face(p)
    assert f2f.shape == (6314,) and set(np.unique(f2f)) <= {1, 2, 3, 4}


def test_halo_run_compression_roundtrip(fv):
    """distributed.runs_of / cols_of: what the ranks actually exchange for their halo lists."""
    import importlib
    d = importlib.import_module("fvb200.distributed")
    rng = np.random.default_rng(11)
    cases = [np.empty(0, np.int64), np.array([5]), np.arange(100, 100 + 262144), np.array([1, 2, 3, 7, 8, 20]),
             np.unique(rng.integers(0, 5000, 1500)),
             np.concatenate([np.arange(10, 500), np.arange(900, 1400)])]
    for c in cases:
        r = d.runs_of(c)
        assert np.array_equal(d.cols_of(r), np.asarray(c, np.int64))
        assert all(n >= 1 for _, n in r) and all(r[i][0] + r[i][1] < r[i + 1][0] for i in range(len(r) - 1))
    assert d.runs_of(np.arange(100, 100 + 262144)) == [(100, 262144)]


@pytest.mark.parametrize("case,world", [("box", 2), ("box", 3), ("box", 4), ("box", 8), ("fractures", 3), ("fractures", 5)])
def test_halo_plan_drives_a_correct_distributed_spmv(fv, orc, fourfractures, case, world):
    """The planner end to end, on the CPU, for rank counts the single-GPU box cannot run: split the oracle's matrix
    into contiguous row ranges (x-slabs of a regular grid; arbitrary ranges of the irregular fracture graph), let
    every rank list the off-rank columns it references, run halo_plan_from_ranges + send_destinations for every rank
    and emulate exactly what the library does per product -- gather the send rows, store them at the destination
    index of the peer's [owned | halo] vector, multiply the local rows -- and compare with the global product."""
    import importlib
    d = importlib.import_module("fvb200.distributed")
    if case == "box":
        ns = [17, 6, 5]
        _, nb, aol, _ = orc.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
        N, plane = int(np.prod(ns)), ns[1] * ns[2]
        k = orc.nodehycos2neighborhycos(nb, np.random.default_rng(0).standard_normal(N), True)
        dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
        A = orc.assembleA(nb, aol, k, np.zeros(N), dn, np.zeros(dn.size), None, True).toscipy().tocsr()
        planes = d.slab_planes(ns[0], world)
        fn, n2f = orc.getfreenodes(N, dn)
        bounds = []
        for lo_p, hi_p in planes:  # free rows of the slab's node range, 0-based [start, end)
            lo, hi = d.node_range_of_planes((lo_p, hi_p), ns[1], ns[2])
            bounds.append((int(np.count_nonzero(fn[:lo - 1])), int(np.count_nonzero(fn[:hi]))))
    else:
        m = fourfractures
        A = orc.assembleA(m["neighbors"], m["areasoverlengths"], m["conductivities"], np.zeros(m["xs"].size),
                          m["dirichletnodes"], m["dirichletheads"]).toscipy().tocsr()
        cuts = np.linspace(0, A.shape[0], world + 1).astype(int)
        bounds = [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]
    nf = A.shape[0]
    assert bounds[0][0] == 0 and bounds[-1][1] == nf and all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    ranges = [(b[0] + 1, b[1] - b[0]) for b in bounds]  # (1-based row_start, nf_local)
    # what fvb_get_halo_cols returns: ascending distinct off-rank columns (1-based global) referenced by the rank's rows
    halos = []
    for lo, hi in bounds:
        cols = np.unique(A[lo:hi].indices)
        halos.append(cols[(cols < lo) | (cols >= hi)].astype(np.int64) + 1)
    plans = [d.halo_plan_from_ranges(r, ranges, halos) for r in range(world)]
    dests = [d.send_destinations(r, plans[r][0], ranges, halos) for r in range(world)]
    x = np.random.default_rng(5).standard_normal(nf)
    vec = [np.concatenate([x[lo:hi], np.full(halos[r].size, np.nan)]) for r, (lo, hi) in enumerate(bounds)]
    for r in range(world):  # k_halo_push of every rank
        peers, send_counts, send_rows, recv_counts = plans[r]
        off = 0
        for p, cnt, dst in zip(peers, send_counts, dests[r]):
            vec[p][dst:dst + cnt] = vec[r][np.asarray(send_rows[off:off + cnt])]
            off += cnt
        assert sum(recv_counts) == halos[r].size
    y = np.empty(nf)
    for r, (lo, hi) in enumerate(bounds):
        assert not np.isnan(vec[r]).any()                          # every halo slot was filled exactly by its owner
        assert np.array_equal(vec[r][hi - lo:], x[halos[r] - 1])    # ... with the right value, in ascending column order
        loc = A[lo:hi].tocoo()
        colmap = np.where((loc.col >= lo) & (loc.col < hi), loc.col - lo,
                          (hi - lo) + np.searchsorted(halos[r] - 1, loc.col))
        yl = np.zeros(hi - lo)
        np.add.at(yl, loc.row, loc.data * vec[r][colmap])
        y[lo:hi] = yl
    assert np.allclose(y, A @ x, rtol=1e-12, atol=1e-14 * np.abs(A).max() * np.abs(x).max())


def _dia_rank_emulation(A, lo, hi, halo, offs, s_global, x_global):
    """numpy transcription of the per-rank index logic of csrc/dia.cuh for rows [lo, hi) of the global matrix A:
    k_dia_fill (U_k[o+r] = A[r, r+o]; U_k[r] = A[r, r-o] only when the partner row is not owned), DiaDesc::xindex
    (owned range, else the low / high halo run), k_dia_scale (S = U * (s_row * s_col)) and the row sum of
    k_spmv_dia<UNIT> on the scaled copy.  `halo` = ascending global columns referenced off-rank."""
    n = hi - lo
    halo = np.asarray(halo, np.int64)
    lo_run, hi_run = halo[halo < lo], halo[halo >= hi]
    assert (lo_run.size == 0 or np.array_equal(lo_run, np.arange(lo_run[0], lo_run[0] + lo_run.size)))
    assert (hi_run.size == 0 or np.array_equal(hi_run, np.arange(hi_run[0], hi_run[0] + hi_run.size)))
    lo0 = int(lo_run[0]) if lo_run.size else 0
    hi0 = int(hi_run[0]) if hi_run.size else 0
    nlo = lo_run.size

    def xindex(g):
        l = g - lo
        if 0 <= l < n:
            return l
        return n + (g - lo0) if g < lo else n + nlo + (g - hi0)

    Ad = A.toarray()
    U = [np.zeros(n + o) for o in offs]
    for r in range(n):                                  # k_dia_fill
        for k, o in enumerate(offs):
            g = lo + r
            if g + o < A.shape[0] and Ad[g, g + o] != 0.0:
                U[k][o + r] = Ad[g, g + o]
            if r < o and g - o >= 0 and Ad[g, g - o] != 0.0:
                U[k][r] = Ad[g, g - o]
    vec_s = np.concatenate([s_global[lo:hi], s_global[halo]])      # [owned | halo] after the halo exchange of s
    vec_x = np.concatenate([x_global[lo:hi], x_global[halo]])
    S = [np.zeros_like(u) for u in U]
    for r in range(n):                                  # k_dia_scale
        for k, o in enumerate(offs):
            up = U[k][o + r]
            if up != 0.0:
                iu = r + o
                S[k][o + r] = up * (vec_s[r] * vec_s[iu if iu < n else xindex(lo + iu)])
            if r < o:
                l = U[k][r]
                S[k][r] = l * (vec_s[r] * vec_s[xindex(lo + r - o)]) if l != 0.0 else 0.0
    y = np.zeros(n)
    for r in range(n):                                  # k_spmv_dia<.., UNIT = true>
        acc = 0.0
        for k in range(len(offs) - 1, -1, -1):
            o, l = offs[k], S[k][r]
            if l != 0.0:
                il = r - o
                acc += l * (vec_x[il] if il >= 0 else vec_x[xindex(lo + il)])
        acc += vec_x[r]
        for k, o in enumerate(offs):
            up = S[k][o + r]
            if up != 0.0:
                iu = r + o
                acc += up * (vec_x[iu] if iu < n else vec_x[xindex(lo + iu)])
        y[r] = acc
    return y


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_scaled_diagonal_format_across_slabs(fv, orc, world):
    """The multi-rank layout of the Jacobi-scaled diagonal copy (symmetric storage with the partner's entries of the
    first plane kept locally, halo runs below and above, scale factors of halo columns taken from the halo slots)
    transcribed to numpy: every rank's rows of D^-1/2 A D^-1/2 x -- including middle ranks with neighbours on both
    sides, which a 2-GPU box cannot exercise -- must equal the global product."""
    import importlib
    d = importlib.import_module("fvb200.distributed")
    import scipy.sparse as sp
    ns = [11, 4, 3]
    _, nb, aol, _ = orc.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
    N, plane = int(np.prod(ns)), ns[1] * ns[2]
    k = orc.nodehycos2neighborhycos(nb, np.random.default_rng(2).standard_normal(N), True)
    dn = np.concatenate([np.arange(1, plane + 1), np.arange(N - plane + 1, N + 1)])
    A = orc.assembleA(nb, aol, k, np.zeros(N), dn, np.zeros(dn.size), None, True).toscipy().tocsr()
    nf = A.shape[0]
    offs = [1, ns[2], plane]
    s = 1.0 / np.sqrt(A.diagonal())
    x = np.random.default_rng(3).standard_normal(nf)
    want = (sp.diags(s) @ A @ sp.diags(s)) @ x
    fn, _ = orc.getfreenodes(N, dn)
    got = np.empty(nf)
    for lo_p, hi_p in d.slab_planes(ns[0], world):
        nlo, nhi = d.node_range_of_planes((lo_p, hi_p), ns[1], ns[2])
        lo, hi = int(np.count_nonzero(fn[:nlo - 1])), int(np.count_nonzero(fn[:nhi]))
        cols = np.unique(A[lo:hi].indices)
        halo = cols[(cols < lo) | (cols >= hi)]
        got[lo:hi] = _dia_rank_emulation(A, lo, hi, halo, offs, s, x)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-15 * np.abs(want).max())


def test_cpp_partition_planning_matches_python(fv, tmp_path):
    """The single-process multi-GPU front end plans its partition in C++ (csrc/host_util.h: slab_planes,
    regulargrid_faces_before, plan_halo_exchange); the one-process-per-GPU mode plans it in Python
    (distributed.slab_planes, grid closed form, halo_plan_from_ranges + send_destinations).  Same answers on slab
    partitions, random irregular partitions and partitions with ranks that own no row."""
    import importlib
    import shutil
    import subprocess
    d = importlib.import_module("fvb200.distributed")
    gxx = shutil.which("g++") or "/usr/bin/g++"
    exe = str(tmp_path / "plan_cli")
    subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "plan_cli.cpp")], check=True)

    def run(*args, stdin=None):
        return subprocess.run([exe, *map(str, args)], input=stdin, capture_output=True, text=True, check=True).stdout

    for n1, P in [(512, 8), (512, 1), (9, 8), (10, 8), (3, 2), (1024, 4), (7, 3), (8, 8)]:
        for ends in (True, False):
            got = [tuple(int(v) for v in ln.split()) for ln in run("planes", n1, P, int(ends)).splitlines()]
            assert got == d.slab_planes(n1, P, ends), (n1, P, ends)
    # closed-form face offsets against the list regulargrid builds
    ns = [4, 3, 5]
    _, nb, _, _ = fv.regulargrid([0, 0, 0], [n - 1 for n in ns], ns, want_coords=False)
    first = {}
    for j, a in enumerate(nb[:, 0]):
        first.setdefault(int(a), j)
    for i1 in range(1, ns[0] + 1):
        for i2 in range(1, ns[1] + 1):
            for i3 in range(1, ns[2] + 1):
                lin = i3 + ns[2] * (i2 - 1) + ns[1] * ns[2] * (i1 - 1)
                if lin in first:
                    assert int(run("faces_before", *ns, i1, i2, i3)) == first[lin]
    # halo plans
    rng = np.random.default_rng(11)
    cases = []
    plane = 12
    for P in (2, 3, 8):  # slabs: one plane of the neighbour on each side
        nf = [plane * int(rng.integers(1, 4)) for _ in range(P)]
        start = np.concatenate([[0], np.cumsum(nf)[:-1]]).tolist()
        halos = []
        for r in range(P):
            h = []
            if r > 0:
                h += list(range(start[r] - plane, start[r]))
            if r + 1 < P:
                h += list(range(start[r] + nf[r], start[r] + nf[r] + plane))
            halos.append(h)
        cases.append((start, nf, halos))
    for _ in range(20):  # irregular: random references, some ranks without rows
        P = int(rng.integers(2, 7))
        nf = [int(rng.integers(0, 9)) for _ in range(P)]
        if sum(nf) == 0:
            nf[0] = 3
        start = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(int).tolist()
        total = sum(nf)
        halos = []
        for r in range(P):
            other = [g for g in range(total) if not (start[r] <= g < start[r] + nf[r])]
            k = int(rng.integers(0, len(other) + 1)) if other else 0
            halos.append(sorted(rng.choice(other, size=k, replace=False).tolist()) if k else [])
        cases.append((start, nf, halos))
    for start, nf, halos in cases:
        P = len(nf)
        text = f"{P}\n" + "".join(f"{start[r]} {nf[r]} {len(halos[r])} " + " ".join(map(str, halos[r])) + "\n" for r in range(P))
        lines = run("halo", stdin=text).splitlines()
        assert not lines[0].startswith("error"), lines[0]
        ranges = [(start[r] + 1, nf[r]) for r in range(P)]  # the Python planner takes 1-based starts and columns
        halos1 = [np.asarray(h, np.int64) + 1 for h in halos]
        for r in range(P):
            blk = {ln.split()[0]: [int(v) for v in ln.split()[1:]] for ln in lines[5 * r:5 * r + 5]}
            peers, sc, sr, rc = d.halo_plan_from_ranges(r, ranges, halos1)
            assert blk["peers"] == list(peers) and blk["send_counts"] == list(sc) and blk["recv_counts"] == list(rc)
            assert blk["send_rows"] == [int(v) for v in sr]
            assert blk["send_dst"] == d.send_destinations(r, peers, ranges, halos1)


def test_bench_implicit_inputs_describe_the_same_workload(fv):
    """bench.py --implicit (the 1024^3 runs) draws the node field in chunks and never builds a face list; it must be the
    very field problem_inputs() draws in one go, slab by slab, so that both input forms time the same workload."""
    import bench
    n = 12
    whole = bench.problem_inputs(fv, n, 1.0)
    plane = n * n
    for planes in [(1, n), (1, 4), (5, 9), (10, 12)]:
        exp = bench.problem_inputs(fv, n, 1.0, planes=planes)
        imp = bench.implicit_inputs(n, 1.0, planes, lambda shape, dt: np.empty(shape, dt))
        assert imp["node_range"] == exp["node_range"] and imp["lnk_node_lo"] == exp["lnk_node_lo"]
        assert np.array_equal(imp["lnk_slab"], exp["lnk_slab"])
        assert np.array_equal(imp["dn"], exp["dn"]) and np.array_equal(imp["dh"], exp["dh"])
        lo = exp["lnk_node_lo"]
        assert np.array_equal(exp["lnk_slab"], whole["lnk_slab"][lo - 1:lo - 1 + exp["lnk_slab"].size])
    # chunked drawing: force several chunks
    big = bench.implicit_inputs(40, 1.0, (3, 20), lambda shape, dt: np.empty(shape, dt), chunk=1000)
    ref = math.log(1e-5) + np.random.default_rng(0).standard_normal(40 ** 3)
    assert np.array_equal(big["lnk_slab"], ref[big["lnk_node_lo"] - 1:big["lnk_node_lo"] - 1 + big["lnk_slab"].size])
