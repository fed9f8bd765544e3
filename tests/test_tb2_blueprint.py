"""CPU blueprint of temporal blocking (oracle/prototypes/tb2_host.cpp, round-2 preparation): one
two-update pass over tiles must equal two single steps of the oracle bit for bit, for any tile shape,
with every boundary rule, the corner quirks and the obstacle inside / across tiles."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "prototypes", "libtb2.so")

CASES = {
    "64x32": O.Case(nx=64, ny=32, cylinder_x=0.3, cylinder_radius=0.2, output_frequency=3),
    "70x33": O.Case(nx=70, ny=33, cylinder_x=0.3, cylinder_radius=0.15, output_frequency=5, inlet_velocity=0.05),
    "cyl_on_wall": O.Case(nx=80, ny=40, cylinder_x=0.5, cylinder_y=0.1, cylinder_radius=0.2, output_frequency=4),
    "cyl_at_inlet": O.Case(nx=72, ny=36, cylinder_x=0.03, cylinder_y=0.5, cylinder_radius=0.25, output_frequency=5),
    "tiny_6x4": O.Case(nx=6, ny=4, cylinder_radius=0.3, output_frequency=2),
}


@pytest.fixture(scope="module")
def tb2():
    L = C.CDLL(LIB)
    L.tb2_pass.restype = C.c_longlong
    L.tb2_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                           C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tile", [(8, 16), (5, 7), (1, 1), (1000, 1000), (16, 3)])
def test_two_updates_per_pass_equal_two_oracle_steps(tb2, name, tile):
    case = CASES[name]
    o = O.Oracle(case)
    for warm in (1, 6):  # start from a post-collision state whose ghost ring is in its permanent form (F4)
        o.run(warm)
        src = np.ascontiguousarray(o.f_next.copy())
        solid = np.ascontiguousarray(o.solid.astype(np.uint8))
        o.run(2)
        want = o.f_next
        dst = np.full_like(src, np.nan)
        b1, b2 = C.c_int(), C.c_int()
        n1 = tb2.tb2_pass(src.ctypes.data, dst.ctypes.data, solid.ctypes.data, case.nx, case.ny, case.tau, case.inlet_velocity,
                          tile[0], tile[1], C.byref(b1), C.byref(b2))
        assert np.array_equal(dst, want), (name, tile, warm, int((dst != want).sum()))
        assert b1.value == 0 and b2.value == 0
        # redundancy of the scheme: stage-1 cells per interior cell
        tx, ty = min(tile[0], case.nx), min(tile[1], case.ny)
        assert n1 >= case.nx * case.ny and n1 <= case.nx * case.ny * (tx + 2) * (ty + 2) / (tx * ty) + 1


def test_redundancy_of_the_planned_gpu_tile(tb2):
    """8 x 128 tiles (the shape planned for shared memory: 9*10*130*8 B = 94 KB): 1.27 stage-1 updates per cell."""
    case = O.Case(nx=64, ny=256, cylinder_radius=0.1, output_frequency=50)
    o = O.Oracle(case)
    o.run(1)
    src = np.ascontiguousarray(o.f_next.copy())
    dst = np.empty_like(src)
    solid = np.ascontiguousarray(o.solid.astype(np.uint8))  # keep alive across the call
    b1, b2 = C.c_int(), C.c_int()
    n1 = tb2.tb2_pass(src.ctypes.data, dst.ctypes.data, solid.ctypes.data, case.nx, case.ny,
                      case.tau, case.inlet_velocity, 8, 128, C.byref(b1), C.byref(b2))
    assert 1.2 < n1 / (case.nx * case.ny) < 1.3
    o.run(2)
    assert np.array_equal(dst, o.f_next)
