"""ctypes view of oracle/liboracle.so and of the oracle/_ref binaries.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does (tests/test_boundary.py greps for that).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_FAST = os.path.join(REF_DIR, "lbm_ref_fast")
REF_STRICT = os.path.join(REF_DIR, "lbm_ref_strict")


def build(quiet: bool = True) -> None:
    """make -C oracle (liboracle.so always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


class _Params(C.Structure):
    _fields_ = [
        ("tau", C.c_double),
        ("inlet_velocity", C.c_double),
        ("gnx", C.c_int),
        ("ny", C.c_int),
        ("output_frequency", C.c_int),
        ("cylinder_x", C.c_double),
        ("cylinder_y", C.c_double),
        ("cylinder_radius", C.c_double),
        ("x_start", C.c_int),
        ("lnx", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        D = C.POINTER(C.c_double)
        L.oracle_create.restype = P
        L.oracle_create.argtypes = [C.POINTER(_Params)]
        L.oracle_destroy.argtypes = [P]
        L.oracle_initialise.argtypes = [P]
        L.oracle_initialise.restype = C.c_int
        for name in ("oracle_collide", "oracle_edge_ghosts", "oracle_stream", "oracle_boundaries"):
            getattr(L, name).argtypes = [P]
            getattr(L, name).restype = None
        L.oracle_forces.argtypes = [P, D, D]
        L.oracle_get_halo.argtypes = [P, C.c_int, D]
        L.oracle_put_halo.argtypes = [P, C.c_int, D]
        L.oracle_check_stability.argtypes = [P]
        L.oracle_check_stability.restype = C.c_int
        L.oracle_max_velocity.argtypes = [P]
        L.oracle_max_velocity.restype = C.c_double
        L.oracle_run.argtypes = [P, C.c_int, C.c_int, D, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_run.restype = C.c_int
        for name in ("oracle_f_current", "oracle_f_next", "oracle_rho", "oracle_ux", "oracle_uy"):
            getattr(L, name).argtypes = [P]
            getattr(L, name).restype = D
        L.oracle_solid.argtypes = [P]
        L.oracle_solid.restype = C.POINTER(C.c_ubyte)
        L.oracle_f_count.argtypes = [P]
        L.oracle_f_count.restype = C.c_size_t
        _lib = L
    return _lib


@dataclass
class Case:
    """The fields of LBM::SimulationParams (reference include/LBMConfig.h:36-52) a case can set."""

    nx: int = 2048
    ny: int = 512
    tau: float = 0.6
    inlet_velocity: float = 0.01333
    output_frequency: int = 140
    cylinder_x: float = 0.2
    cylinder_y: float = 0.5
    cylinder_radius: float = 0.05

    def ref_args(self, steps: int):
        return [
            "--nx", str(self.nx), "--ny", str(self.ny), "--steps", str(steps), "--of", str(self.output_frequency),
            "--tau", repr(self.tau), "--uin", repr(self.inlet_velocity), "--cx", repr(self.cylinder_x),
            "--cy", repr(self.cylinder_y), "--cr", repr(self.cylinder_radius),
        ]


class Oracle:
    """One slab of the CPU restatement.  x_start=0, lnx=nx is the reference's 1-rank job."""

    def __init__(self, case: Case, x_start: int = 0, lnx: int | None = None):
        self.case = case
        self.lnx = case.nx if lnx is None else lnx
        self.x_start = x_start
        p = _Params(case.tau, case.inlet_velocity, case.nx, case.ny, case.output_frequency, case.cylinder_x,
                    case.cylinder_y, case.cylinder_radius, x_start, self.lnx)
        self._h = lib().oracle_create(C.byref(p))
        if not self._h:
            raise MemoryError("oracle_create failed")
        self.n_solid = lib().oracle_initialise(self._h)
        self.t = 0

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # --- views on the C arrays (no copies) ---------------------------------------------
    def _view(self, fn, n, dtype=np.float64):
        ptr = fn(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n,)) if dtype == np.float64 else np.ctypeslib.as_array(ptr, shape=(n,))

    @property
    def f_current(self):
        return self._view(lib().oracle_f_current, lib().oracle_f_count(self._h)).reshape(self.case.ny + 2, self.lnx + 2, 9)

    @property
    def f_next(self):
        return self._view(lib().oracle_f_next, lib().oracle_f_count(self._h)).reshape(self.case.ny + 2, self.lnx + 2, 9)

    @property
    def rho(self):
        return self._view(lib().oracle_rho, self.lnx * self.case.ny).reshape(self.case.ny, self.lnx)

    @property
    def ux(self):
        return self._view(lib().oracle_ux, self.lnx * self.case.ny).reshape(self.case.ny, self.lnx)

    @property
    def uy(self):
        return self._view(lib().oracle_uy, self.lnx * self.case.ny).reshape(self.case.ny, self.lnx)

    @property
    def solid(self):
        return self._view(lib().oracle_solid, self.lnx * self.case.ny, np.uint8).reshape(self.case.ny, self.lnx)

    # --- phases ------------------------------------------------------------------------
    def collide(self):
        lib().oracle_collide(self._h)

    def forces(self):
        fx, fy = C.c_double(), C.c_double()
        lib().oracle_forces(self._h, C.byref(fx), C.byref(fy))
        return fx.value, fy.value

    def edge_ghosts(self):
        lib().oracle_edge_ghosts(self._h)

    def get_halo(self, east: bool):
        buf = np.empty(9 * self.case.ny, dtype=np.float64)
        lib().oracle_get_halo(self._h, int(east), buf.ctypes.data_as(C.POINTER(C.c_double)))
        return buf

    def put_halo(self, east: bool, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        lib().oracle_put_halo(self._h, int(east), buf.ctypes.data_as(C.POINTER(C.c_double)))

    def stream(self):
        lib().oracle_stream(self._h)

    def boundaries(self):
        lib().oracle_boundaries(self._h)

    def check_stability(self) -> bool:
        return bool(lib().oracle_check_stability(self._h))

    def max_velocity(self) -> float:
        return lib().oracle_max_velocity(self._h)

    def run(self, nsteps: int):
        """Solver::run for nsteps more iterations.  Returns (forces rows [n,5], unstable_at or -1)."""
        max_rows = nsteps // max(self.case.output_frequency, 1) + 2
        rows = np.zeros((max_rows, 5), dtype=np.float64)
        n_rows, bad = C.c_int(), C.c_int()
        done = lib().oracle_run(self._h, self.t, nsteps, rows.ctypes.data_as(C.POINTER(C.c_double)), max_rows,
                                C.byref(n_rows), C.byref(bad))
        self.t += done
        return rows[: n_rows.value].copy(), bad.value


def format_forces_csv(rows) -> str:
    """forces.csv exactly as IOManager writes it (reference include/LBMIO.h:40,180-185)."""
    out = ["timestep,drag_force,lift_force,drag_coeff,lift_coeff\n"]
    for r in rows:
        out.append("%d,%.8f,%.8f,%.8f,%.8f\n" % (int(r[0]), r[1], r[2], r[3], r[4]))
    return "".join(out)


def have_ref() -> bool:
    return os.path.exists(REF_FAST) and os.path.exists(REF_STRICT)


def run_ref(case: Case, steps: int, strict: bool = True, threads: int = 1):
    """Run the compiled reference (oracle/_ref) and return its observable state as a dict."""
    exe = REF_STRICT if strict else REF_FAST
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        r = subprocess.run([exe] + case.ref_args(steps) + ["--dump", d], env=env, capture_output=True, text=True)
        ny, nx = case.ny, case.nx
        out = {
            "f_current": np.fromfile(os.path.join(d, "f_current.bin")).reshape(ny + 2, nx + 2, 9),
            "f_next": np.fromfile(os.path.join(d, "f_next.bin")).reshape(ny + 2, nx + 2, 9),
            "rho": np.fromfile(os.path.join(d, "rho.bin")).reshape(ny, nx),
            "ux": np.fromfile(os.path.join(d, "ux.bin")).reshape(ny, nx),
            "uy": np.fromfile(os.path.join(d, "uy.bin")).reshape(ny, nx),
            "solid": np.fromfile(os.path.join(d, "solid.bin"), dtype=np.uint8).reshape(ny, nx),
            "forces_csv": open(os.path.join(d, "forces.csv")).read(),
            "stdout": r.stdout,
            "stderr": r.stderr,
            "returncode": r.returncode,
        }
    return out


def time_ref(case: Case, steps: int, warmup: int, threads: int, timeout: float = 1200.0):
    """Time Solver::run of the fast-math reference build; returns the harness's JSON dict."""
    import json

    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="true")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([REF_FAST] + case.ref_args(steps) + ["--time", "--warmup", str(warmup)], env=env, cwd=d,
                           capture_output=True, text=True, timeout=timeout)
    for line in reversed(r.stderr.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("reference timing failed: rc=%d\n%s" % (r.returncode, r.stderr[-2000:]))
