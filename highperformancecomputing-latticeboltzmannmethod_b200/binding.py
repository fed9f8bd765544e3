"""ctypes binding of liblbm_b200.so (include/lbm_b200.h) for tests/ and bench.py.

This is plumbing, not product: the product is the C-ABI library plus the C++ drop-in headers in
include/ (LBMConfig.h / LBMGrid.h / LBMSolver.h / LBMIO.h).  There is no CPU fallback: if the
library is missing or no CUDA device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "liblbm_b200.so")

FLAG_PERIODIC_X = 1
FLAG_PERIODIC_Y = 2
FLAG_NO_CYLINDER = 4
FLAG_SHEAR_WAVE_INIT = 8
FLAG_AA = 16

F_CURRENT, F_NEXT = 0, 1
VARIANT_SCALAR, VARIANT_VEC2, VARIANT_TB = 0, 1, 2


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("liblbm_b200 error %d: %s" % (code, msg))
        self.code = code


class CParams(C.Structure):
    """struct lbm_params -- LBM::SimulationParams (reference include/LBMConfig.h:36-52) + extensions."""

    _fields_ = [
        ("tau", C.c_double),
        ("inlet_velocity", C.c_double),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("num_timesteps", C.c_int32),
        ("output_frequency", C.c_int32),
        ("cylinder_x", C.c_double),
        ("cylinder_y", C.c_double),
        ("cylinder_radius", C.c_double),
        ("vtk_start_step", C.c_int32),
        ("flags", C.c_int32),
        ("body_force_x", C.c_double),
        ("body_force_y", C.c_double),
    ]


class CInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "global_nx", "global_ny", "local_nx", "local_ny", "x_start", "y_start", "rank", "world",
        "device", "cyl_x", "cyl_y", "cyl_r", "solid_cells", "links", "iteration")] + [
        ("bytes_per_buffer", C.c_int64), ("row_pitch", C.c_int32), ("kernel_variant", C.c_int32), ("halo_p2p", C.c_int32),
        ("pass_depth", C.c_int32), ("deep_solid_cells", C.c_int32)]


@dataclass
class SimulationParams:
    """Same fields, defaults and derived quantities as LBM::SimulationParams
    (reference include/LBMConfig.h:36-66), plus the extension fields of lbm_params."""

    tau: float = 0.6
    inlet_velocity: float = 0.01333
    nx: int = 2048
    ny: int = 512
    num_timesteps: int = 120000
    output_frequency: int = 140
    cylinder_x: float = 0.2
    cylinder_y: float = 0.5
    cylinder_radius: float = 0.05
    vtk_start_step: int = 0
    flags: int = 0
    body_force_x: float = 0.0
    body_force_y: float = 0.0

    def nu(self):
        return (self.tau - 0.5) / 3.0

    def reynolds(self):
        return self.inlet_velocity * (2.0 * self.cylinder_radius * self.ny) / self.nu()

    def get_cylinder_x(self):
        return int(self.cylinder_x * self.nx)

    def get_cylinder_y(self):
        return int(self.cylinder_y * self.ny)

    def get_cylinder_radius_cells(self):
        return int(self.cylinder_radius * self.ny)

    def to_c(self) -> CParams:
        return CParams(self.tau, self.inlet_velocity, self.nx, self.ny, self.num_timesteps, self.output_frequency,
                       self.cylinder_x, self.cylinder_y, self.cylinder_radius, self.vtk_start_step, self.flags,
                       self.body_force_x, self.body_force_y)


EXPORTS = [
    "lbm_create", "lbm_create_slab", "lbm_nccl_unique_id", "lbm_destroy", "lbm_last_error", "lbm_get_info",
    "lbm_setup_geometry", "lbm_initialise", "lbm_step", "lbm_run", "lbm_sync", "lbm_get_forces",
    "lbm_check_stability", "lbm_max_velocity", "lbm_download_f", "lbm_download_macros", "lbm_download_solid",
    "lbm_upload_f", "lbm_snapshot_begin", "lbm_snapshot_wait", "lbm_host_alloc", "lbm_host_free", "lbm_time_steps",
    "lbm_set_kernel_variant", "lbm_device_count", "lbm_get_counters", "lbm_event_record", "lbm_event_elapsed",
    "lbm_bootstrap_env", "lbm_set_params", "lbm_snapshot_begin_slot", "lbm_snapshot_wait_slot", "lbm_allreduce", "lbm_gather_macros",
    "lbm_get_bulk_updates", "lbm_set_pass_depth", "lbm_set_force_mode", "lbm_snapshot_begin_slot2d", "lbm_host_register",
    "lbm_host_unregister", "lbm_upload_f_next", "lbm_plan_passes", "lbm_selftest_division",
]

_lib = None


def load():
    """Load the C-ABI library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a fresh checkout: compile the sm_100a library in place (nvcc needs no GPU); still no library -> fail loudly
        import subprocess

        subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j4"], capture_output=True)
    if not os.path.exists(LIB_PATH):
        raise LbmError(-2, "liblbm_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    H, D, I = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.lbm_create.argtypes = [C.POINTER(CParams), C.c_int, C.POINTER(H)]
    L.lbm_create_slab.argtypes = [C.POINTER(CParams), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(H)]
    L.lbm_nccl_unique_id.argtypes = [C.c_void_p]
    L.lbm_destroy.argtypes = [H]
    L.lbm_last_error.argtypes = [H]
    L.lbm_last_error.restype = C.c_char_p
    L.lbm_get_info.argtypes = [H, C.POINTER(CInfo)]
    L.lbm_setup_geometry.argtypes = [H, I]
    L.lbm_initialise.argtypes = [H, C.c_double]
    L.lbm_step.argtypes = [H, C.c_int]
    L.lbm_run.argtypes = [H, C.c_int, D, C.c_int, I, I]
    L.lbm_sync.argtypes = [H]
    L.lbm_get_forces.argtypes = [H, D, D]
    L.lbm_check_stability.argtypes = [H, I, I]
    L.lbm_max_velocity.argtypes = [H, D]
    L.lbm_download_f.argtypes = [H, C.c_int, C.c_void_p]
    L.lbm_download_macros.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbm_download_solid.argtypes = [H, C.c_void_p]
    L.lbm_upload_f.argtypes = [H, C.c_void_p, C.c_int]
    L.lbm_snapshot_begin.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbm_snapshot_wait.argtypes = [H]
    L.lbm_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.lbm_host_free.argtypes = [C.c_void_p]
    L.lbm_time_steps.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), I]
    L.lbm_set_kernel_variant.argtypes = [H, C.c_int]
    L.lbm_device_count.argtypes = [I]
    LL = C.POINTER(C.c_longlong)
    L.lbm_get_counters.argtypes = [H, LL, LL, LL]
    L.lbm_event_record.argtypes = [H, C.c_int]
    L.lbm_event_elapsed.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.lbm_snapshot_begin_slot.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbm_snapshot_wait_slot.argtypes = [H, C.c_int]
    L.lbm_bootstrap_env.argtypes = [I, I, I, C.c_void_p]
    L.lbm_set_params.argtypes = [H, C.POINTER(CParams)]
    L.lbm_allreduce.argtypes = [H, D, C.c_int, C.c_int]
    L.lbm_gather_macros.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lbm_get_bulk_updates.argtypes = [H, LL]
    L.lbm_set_pass_depth.argtypes = [H, C.c_int]
    L.lbm_set_force_mode.argtypes = [H, C.c_int]
    L.lbm_snapshot_begin_slot2d.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.lbm_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.lbm_host_unregister.argtypes = [C.c_void_p]
    L.lbm_upload_f_next.argtypes = [H, C.c_void_p]
    L.lbm_selftest_division.argtypes = [H, C.c_longlong, C.c_ulonglong, LL]
    L.lbm_plan_passes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, I, C.c_int]
    for name in EXPORTS:
        if name != "lbm_last_error":
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().lbm_nccl_unique_id(buf)
    if rc:
        raise LbmError(rc, load().lbm_last_error(None).decode())
    return buf.raw


def pinned_empty(shape, dtype=np.float64):
    """numpy array over cudaHostAlloc'ed memory (lbm_host_alloc).  Keep the returned owner alive."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = load().lbm_host_alloc(C.byref(p), n)
    if rc:
        raise LbmError(rc, load().lbm_last_error(None).decode())
    buf = (C.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def pinned_free(p):
    load().lbm_host_free(p)


@dataclass
class Solver:
    """One handle = one x-slab on one GPU.  Method names follow the reference's Solver / Grid /
    IOManager members they stand for (include/LBMSolver.h, LBMGrid.h, LBMIO.h)."""

    params: SimulationParams
    device: int = 0
    rank: int = 0
    world: int = 1
    nccl_id: bytes | None = None
    _h: C.c_void_p = field(default=None, repr=False)

    def __post_init__(self):
        L = load()
        h = C.c_void_p()
        cp = self.params.to_c()
        if self.world == 1:
            rc = L.lbm_create(C.byref(cp), self.device, C.byref(h))
        else:
            rc = L.lbm_create_slab(C.byref(cp), self.device, self.rank, self.world, self.nccl_id, C.byref(h))
        if rc:
            raise LbmError(rc, L.lbm_last_error(None).decode())
        self._h = h

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise LbmError(rc, load().lbm_last_error(self._h).decode())

    def close(self):
        if self._h:
            load().lbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> CInfo:
        i = CInfo()
        self._ck(load().lbm_get_info(self._h, C.byref(i)))
        return i

    # -- Solver::initialise (LBMSolver.h:31-41) ---------------------------------------------
    def setup_geometry(self) -> int:
        n = C.c_int()
        self._ck(load().lbm_setup_geometry(self._h, C.byref(n)))
        return n.value

    def initialise(self, inlet_u: float | None = None):
        """Solver::initialise: geometry, then Grid::initialise(inlet_velocity)."""
        self.setup_geometry()
        self._ck(load().lbm_initialise(self._h, self.params.inlet_velocity if inlet_u is None else inlet_u))

    # -- Solver::run (LBMSolver.h:43-78) ----------------------------------------------------
    def step(self, n: int = 1):
        self._ck(load().lbm_step(self._h, n))

    def sync(self):
        self._ck(load().lbm_sync(self._h))

    def run(self, n: int):
        """Returns (forces rows [k,5] = t,Fx,Fy,C_D,C_L ; unstable_at or -1)."""
        max_rows = n // max(self.params.output_frequency, 1) + 2
        rows = np.zeros((max_rows, 5))
        k, bad = C.c_int(), C.c_int()
        self._ck(load().lbm_run(self._h, n, rows.ctypes.data_as(C.POINTER(C.c_double)), max_rows, C.byref(k), C.byref(bad)))
        return rows[: k.value].copy(), bad.value

    def forces(self):
        fx, fy = C.c_double(), C.c_double()
        self._ck(load().lbm_get_forces(self._h, C.byref(fx), C.byref(fy)))
        return fx.value, fy.value

    def check_stability(self):
        ok, bad = C.c_int(), C.c_int()
        self._ck(load().lbm_check_stability(self._h, C.byref(ok), C.byref(bad)))
        return bool(ok.value), bad.value

    def max_velocity(self) -> float:
        v = C.c_double()
        self._ck(load().lbm_max_velocity(self._h, C.byref(v)))
        return v.value

    # -- Grid accessors (LBMGrid.h:115-129,145) ---------------------------------------------
    def f(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        i = self.info()
        if out is None:
            out = np.empty((i.local_ny + 2, i.local_nx + 2, 9))
        assert out.shape == (i.local_ny + 2, i.local_nx + 2, 9) and out.flags.c_contiguous and out.dtype == np.float64
        self._ck(load().lbm_download_f(self._h, which, out.ctypes.data))
        return out

    def f_current(self):
        return self.f(F_CURRENT)

    def f_next(self):
        return self.f(F_NEXT)

    def macros(self, out=None):
        i = self.info()
        rho, ux, uy = out if out is not None else (np.empty((i.local_ny, i.local_nx)) for _ in range(3))
        for a in (rho, ux, uy):
            assert a.shape == (i.local_ny, i.local_nx) and a.flags.c_contiguous and a.dtype == np.float64
        self._ck(load().lbm_download_macros(self._h, rho.ctypes.data, ux.ctypes.data, uy.ctypes.data))
        return rho, ux, uy

    def solid(self):
        i = self.info()
        m = np.empty((i.local_ny, i.local_nx), dtype=np.uint8)
        self._ck(load().lbm_download_solid(self._h, m.ctypes.data))
        return m

    def upload_f(self, f_current: np.ndarray, iteration: int = 0):
        a = np.ascontiguousarray(f_current, dtype=np.float64)
        i = self.info()
        assert a.shape == (i.local_ny + 2, i.local_nx + 2, 9), a.shape
        self._ck(load().lbm_upload_f(self._h, a.ctypes.data, iteration))

    def upload_f_next(self, f_next: np.ndarray):
        """Install caller-written f_next values (solid cells, S/N ghost rows, and the fluid cells' newest state)."""
        a = np.ascontiguousarray(f_next, dtype=np.float64)
        i = self.info()
        assert a.shape == (i.local_ny + 2, i.local_nx + 2, 9), a.shape
        self._ck(load().lbm_upload_f_next(self._h, a.ctypes.data))

    def snapshot_begin(self, rho, ux, uy):
        self._ck(load().lbm_snapshot_begin(self._h, rho.ctypes.data, ux.ctypes.data, uy.ctypes.data))

    def snapshot_wait(self):
        self._ck(load().lbm_snapshot_wait(self._h))

    # -- measurement ------------------------------------------------------------------------
    def allreduce(self, values, op: int = 0) -> np.ndarray:
        """In-place NCCL all-reduce of a few host doubles over the slabs (0 sum, 1 min, 2 max)."""
        a = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._ck(load().lbm_allreduce(self._h, a.ctypes.data_as(C.POINTER(C.c_double)), a.size, op))
        return a

    def gather_macros(self):
        """rho, ux, uy of the whole channel on rank 0 ([global_ny, global_nx]); None elsewhere."""
        i = self.info()
        if i.rank == 0:
            out = tuple(np.empty((i.global_ny, i.global_nx)) for _ in range(3))
            self._ck(load().lbm_gather_macros(self._h, *(a.ctypes.data for a in out)))
            return out
        self._ck(load().lbm_gather_macros(self._h, None, None, None))
        return None

    def set_params(self, params: "SimulationParams"):
        cp = params.to_c()
        self._ck(load().lbm_set_params(self._h, C.byref(cp)))
        self.params = params

    def time_steps(self, n: int, per_kernel: int = 0):
        """(ms_total, ms_bulk_kernels, launches) for n iterations, CUDA events on the compute stream."""
        a, b, l = C.c_float(), C.c_float(), C.c_int()
        self._ck(load().lbm_time_steps(self._h, n, int(per_kernel), C.byref(a), C.byref(b), C.byref(l)))
        return a.value, b.value, l.value

    def counters(self):
        """(launches since creation, bulk launches, bulk cells) -- the last two for the last
        time_steps(per_kernel=True)."""
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self._ck(load().lbm_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def event_record(self, slot: int):
        self._ck(load().lbm_event_record(self._h, slot))

    def event_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._ck(load().lbm_event_elapsed(self._h, a, b, C.byref(ms)))
        return ms.value

    def set_kernel_variant(self, v: int):
        self._ck(load().lbm_set_kernel_variant(self._h, v))

    def set_pass_depth(self, d: int):
        """Iterations per temporally blocked pass (kernel variant 2): 1, 2 (default) or 3."""
        self._ck(load().lbm_set_pass_depth(self._h, d))

    def selftest_division(self, n: int, seed: int = 1) -> int:
        """Quotients of div_pair (the kernels' u = j / rho) that differ from the IEEE division on n random triples."""
        bad = C.c_longlong()
        self._ck(load().lbm_selftest_division(self._h, n, seed, C.byref(bad)))
        return bad.value

    def set_force_mode(self, tree: int):
        """0: the reference's serial summation order (its bits); 1: fixed parallel tree (a few us, equal to rounding)."""
        self._ck(load().lbm_set_force_mode(self._h, tree))

    def bulk_updates(self) -> int:
        """Cell updates of the bulk launches the last time_steps(per_kernel) timed."""
        a = C.c_longlong()
        self._ck(load().lbm_get_bulk_updates(self._h, C.byref(a)))
        return a.value


def plan_passes(iteration: int, n_steps: int, output_frequency: int, max_depth: int = 3, state_is_f_current: bool = False):
    """The pass depths lbm_step(n_steps) launches from `iteration` on (host logic only, no device)."""
    out = (C.c_int * max(n_steps, 1))()
    n = load().lbm_plan_passes(iteration, n_steps, output_frequency, max_depth, 1 if state_is_f_current else 0, out, n_steps)
    if n < 0:
        raise LbmError(n, "bad arguments")
    return list(out[:n])


def device_count() -> int:
    n = C.c_int()
    load().lbm_device_count(C.byref(n))
    return n.value
