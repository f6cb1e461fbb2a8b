"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/vof.h
declares, fails loudly without a GPU (no CPU fallback), and the slab/driver host logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for hdr in ("vof.h",):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(vof[0-9a-z]*_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert len(names) >= 30
    raw = C.CDLL(os.path.join(ROOT, "taichi_2d_vof_b200", "libvof.so"))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in include/vof.h but not exported: {missing}"
    from taichi_2d_vof_b200 import _lib
    assert set(_lib.SIGNATURES) == set(names), set(names) ^ set(_lib.SIGNATURES)


def test_abi_version_and_default_params(built_lib):
    from taichi_2d_vof_b200 import VofParams
    assert built_lib.vof_abi_version() == 1
    p = VofParams()
    built_lib.vof_default_params(C.byref(p))
    assert (p.nx, p.ny, p.n_jacobi) == (200, 200, 10)
    assert (p.rho_l, p.rho_g, p.sigma, p.gy, p.dt) == (1000.0, 50.0, 0.007, -5.0, 4e-6)
    assert built_lib.vof2d_arena_bytes(C.byref(p)) > 12 * 202 * 202 * 4


def test_params_match_oracle_constants():
    from oracle.vof2d_oracle import Vof2DParams
    from taichi_2d_vof_b200 import reference_params, scaled_params
    for n in (200, 2048, 8192):
        a, b = scaled_params(n), Vof2DParams.scaled(n)
        assert a.dx == b.dx and a.dy == b.dy and a.Lx == b.Lx
    assert reference_params().dx == Vof2DParams().dx


def test_no_cpu_fallback(built_lib):
    """Without a usable GPU creation must FAIL (VOF_ENODEV), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from taichi_2d_vof_b200 import VofError, VofSolver2D
    with pytest.raises(VofError) as e:
        VofSolver2D()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_argument_errors(built_lib):
    from taichi_2d_vof_b200 import VofParams
    h = C.c_void_p()
    p = VofParams()
    built_lib.vof_default_params(C.byref(p))
    p.nx = 2
    assert built_lib.vof2d_create(C.byref(p), C.byref(h)) == -1
    assert b"nx, ny" in built_lib.vof_last_error()
    built_lib.vof_default_params(C.byref(p))
    p.slab_lo, p.slab_hi, p.halo = 1, 100, 5          # a slab needs halo >= n_jacobi + 5
    assert built_lib.vof2d_create(C.byref(p), C.byref(h)) == -1
    assert b"halo" in built_lib.vof_last_error()
    assert built_lib.vof2d_set_BC(None) == -1


def test_driver_cli_matches_reference_flags():
    from taichi_2d_vof_b200.driver import build_parser
    p = build_parser()
    a = p.parse_args([])
    assert a.ic == 1 and a.s is False and (a.nx, a.ny, a.jacobi, a.nstep) == (200, 200, 10, 100)
    a = p.parse_args(["-ic", "3", "-s"])
    assert a.ic == 3 and a.s is True
    with pytest.raises(SystemExit):
        p.parse_args(["-ic", "4"])


def test_partition_and_halo_blocks():
    from taichi_2d_vof_b200.slab import halo_row_blocks, partition, required_halo
    assert partition(8192, 1) == [(1, 8192)]
    parts = partition(32768, 8)
    assert parts[0] == (1, 4096) and parts[-1] == (28673, 32768)
    parts = partition(10, 3)
    assert parts == [(1, 4), (5, 7), (8, 10)]
    assert required_halo(10) == 15
    assert halo_row_blocks(100, 16) == ((16, 32), (0, 16), (68, 84), (84, 100))
    with pytest.raises(ValueError):
        partition(2, 3)


def _slab_worker(rank, world, port, H, ny):
    import torch
    import torch.distributed as dist
    from taichi_2d_vof_b200.slab import exchange, halo_row_blocks, partition
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nx = 96
    lo, hi = partition(nx, world)[rank]
    nrows = (hi - lo + 1) + 2 * H
    gi0 = lo - H
    fields = []
    for f in range(4):   # 4 "fields": value = 1000*f + global row, halos poisoned
        t = torch.full((nrows, ny), -1.0)
        t[H:nrows - H] = (torch.arange(lo, hi + 1, dtype=torch.float32) + 1000 * f)[:, None]
        fields.append(t)
    (sa, sb), (ra, rb), (ua, ub), (va, vb) = halo_row_blocks(nrows, H)
    flat = lambda t, a, b: t[a:b].reshape(-1)
    exchange(dist, rank, world, [flat(t, sa, sb) for t in fields], [flat(t, ra, rb) for t in fields],
             [flat(t, ua, ub) for t in fields], [flat(t, va, vb) for t in fields])
    for f, t in enumerate(fields):
        rows = t[:, 0] - 1000 * f
        for l in range(nrows):
            gi = gi0 + l
            if 1 <= gi <= nx:
                assert rows[l] == gi, (rank, f, l, float(rows[l]), gi)      # every real row now holds its global index
            else:
                assert t[l, 0] == -1.0                                      # beyond the physical wall: untouched
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_slab_worker, args=(world, port, 16, 8), nprocs=world, join=True)


def test_python_constants_match_header():
    """Every VOF_OPT_* / VOF_VIEW_* / VOF_STEP_* constant the Python side uses has the value include/vof.h gives it."""
    import re
    from taichi_2d_vof_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vof.h")).read()
    found = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(VOF_(?:OPT|VIEW)_[A-Z_0-9]+)\s*=\s*(\d+)", hdr))
    assert {"VOF_OPT_JACOBI_TB", "VOF_OPT_ADAPTIVE", "VOF_OPT_CHUNK_CAP", "VOF_VIEW_VOF", "VOF_VIEW_VNORM"} <= set(found)
    for name, value in found.items():
        assert hasattr(_lib, name), f"{name} is missing from taichi_2d_vof_b200/_lib.py"
        assert getattr(_lib, name) == value, f"{name}: header {value}, _lib.py {getattr(_lib, name)}"
    assert len(set(v for k, v in found.items() if k.startswith("VOF_OPT_"))) == len([k for k in found if k.startswith("VOF_OPT_")]), "duplicate option ids"


def test_driver3d_cli_matches_reference_flags():
    from taichi_2d_vof_b200.driver3d import build_parser
    p = build_parser()
    a = p.parse_args([])
    assert a.ic == 1 and a.s is False and (a.nx, a.ny, a.nz, a.jacobi, a.nstep) == (200, 200, 200, 10, 100)   # 3dvof.py:12-24, 590
    a = p.parse_args(["-ic", "2", "-s"])
    assert a.ic == 2 and a.s is True
    with pytest.raises(SystemExit):
        p.parse_args(["-ic", "0"])


def test_vtr_export_roundtrip(tmp_path):
    """The .vtr the 3-D driver writes in place of pyevtk.gridToVTK (3dvof.py:624-627): structure and payload."""
    from taichi_2d_vof_b200.vtk import grid_to_vtk, read_vtr_arrays
    rng = np.random.default_rng(0)
    n = (5, 4, 7)
    x, y, z = (np.linspace(0.0, 1.0, k).astype(np.float32) for k in n)
    F = rng.random(n, dtype=np.float32)
    fname = grid_to_vtk(str(tmp_path / "step-00100"), x, y, z, pointData={"VOF": F})
    assert fname.endswith("step-00100.vtr")
    text = open(fname, "rb").read()
    assert b'type="RectilinearGrid"' in text and b'WholeExtent="0 4 0 3 0 6"' in text and b'Name="VOF"' in text
    back = read_vtr_arrays(fname)
    assert back["shape"] == n
    assert np.array_equal(back["VOF"], F)                      # point (i, j, k) -> F[i, j, k], x fastest in the file
    assert np.array_equal(back["x_coordinates"], x) and np.array_equal(back["z_coordinates"], z)
    with pytest.raises(ValueError):
        grid_to_vtk(str(tmp_path / "bad"), x, y, z, pointData={"VOF": F[:-1]})


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the reference arm of the contract: the CPU restatement on the host cores) prints one
    JSON line with the contract's keys; it runs without a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--n", "128"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["dtype"] == "f32" and d["vs_baseline"] is None
