"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref).  TEST INFRASTRUCTURE ONLY.

Run in the build container, where /root/reference exists:

    python oracle/gen_golden.py            # (re)writes tests/golden/

The reference ships no golden vectors, so the pins are outputs of the
reference itself: its unmodified headers compiled by oracle/Makefile and run at
OMP_NUM_THREADS=1 (the multi-threaded boundary loop races at two corner cells).
Two builds are recorded:
  strict  -O2 -ffp-contract=off, no fast-math: bit-reproducible; lbm_oracle.c must equal it
          bit for bit.
  fast    the reference's own flags (-O3 -ffast-math -mfma ...): what a user of the reference
          actually runs; everything must agree with it to <= 1e-12 relative on populations.
Small cases are stored whole, larger ones as sampled cells (the four quirky domain corners, the
inlet and outlet columns, wall rows, the neighbourhood of the cylinder incl. solid cells,
ghost cells, and seeded random cells) plus forces.csv.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(O.HERE), "tests", "golden")

FULL_CASES = {
    # name: (Case, [N...])
    "full_64x32": (O.Case(nx=64, ny=32, cylinder_x=0.3, cylinder_radius=0.2, output_frequency=3), [1, 2, 10, 100]),
}
SAMPLED_CASES = {
    "sampled_256x64": (O.Case(nx=256, ny=64, output_frequency=140), [1, 10, 100, 1000]),
    "sampled_2048x512_default": (O.Case(), [1, 2, 10, 100, 300, 1000]),
}


def sample_coords(case: O.Case, seed: int = 1234):
    """Padded (gx, gy) coordinates to keep: deterministic, covers every special region."""
    nx, ny = case.nx, case.ny
    pts = set()
    # the padded corners + the four interior corner cells and their neighbours
    for gx in (0, 1, 2, nx - 1, nx, nx + 1):
        for gy in (0, 1, 2, ny - 1, ny, ny + 1):
            pts.add((gx, gy))
    # inlet / outlet columns and the ghost columns beside them (strided)
    for gy in range(0, ny + 2, max(1, ny // 32)):
        for gx in (0, 1, 2, nx - 1, nx, nx + 1):
            pts.add((gx, gy))
    # wall rows and ghost rows (strided)
    for gx in range(0, nx + 2, max(1, nx // 64)):
        for gy in (0, 1, 2, ny - 1, ny, ny + 1):
            pts.add((gx, gy))
    # box around the cylinder (solid cells, link cells, wake)
    cx, cy, r = int(case.cylinder_x * nx), int(case.cylinder_y * ny), int(case.cylinder_radius * ny)
    step = max(1, r // 8)
    for y in range(max(0, cy - r - 3), min(ny, cy + r + 4), step):
        for x in range(max(0, cx - r - 3), min(nx, cx + 3 * r + 4), step):
            pts.add((x + 1, y + 1))
    rng = np.random.default_rng(seed)
    for _ in range(512):
        pts.add((int(rng.integers(0, nx + 2)), int(rng.integers(0, ny + 2))))
    arr = np.array(sorted(pts), dtype=np.int32)
    return arr[:, 0], arr[:, 1]


def interior_subset(case, gx, gy):
    m = (gx >= 1) & (gx <= case.nx) & (gy >= 1) & (gy <= case.ny)
    return m


def main():
    if not O.have_ref():
        O.build(quiet=False)
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built (no /root/reference?)")
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {}

    for name, (case, steps_list) in FULL_CASES.items():
        out = {}
        for n in steps_list:
            for build in ("strict", "fast"):
                ref = O.run_ref(case, n, strict=(build == "strict"), threads=1)
                assert ref["returncode"] == 0
                for k in ("f_current", "f_next", "rho", "ux", "uy"):
                    out[f"{build}_N{n}_{k}"] = ref[k]
                out[f"{build}_N{n}_forces_csv"] = np.frombuffer(ref["forces_csv"].encode(), dtype=np.uint8)
                if build == "strict":
                    out["solid"] = ref["solid"]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        manifest[name] = {"case": case.__dict__, "steps": steps_list, "kind": "full"}
        print("wrote", name)

    for name, (case, steps_list) in SAMPLED_CASES.items():
        gx, gy = sample_coords(case)
        m = interior_subset(case, gx, gy)
        out = {"gx": gx, "gy": gy, "interior": m}
        for n in steps_list:
            for build in ("strict", "fast"):
                ref = O.run_ref(case, n, strict=(build == "strict"), threads=1)
                assert ref["returncode"] == 0
                out[f"{build}_N{n}_f_current"] = ref["f_current"][gy, gx, :]
                out[f"{build}_N{n}_f_next"] = ref["f_next"][gy, gx, :]
                for k in ("rho", "ux", "uy"):
                    out[f"{build}_N{n}_{k}"] = ref[k][gy[m] - 1, gx[m] - 1]
                out[f"{build}_N{n}_forces_csv"] = np.frombuffer(ref["forces_csv"].encode(), dtype=np.uint8)
                # whole-field invariants that tolerate a tolerance: total mass and momentum
                fc = ref["f_current"][1:-1, 1:-1, :]
                out[f"{build}_N{n}_sums"] = np.array([fc.sum(), (fc ** 2).sum(), ref["rho"].sum(), ref["ux"].sum()])
                if build == "strict":
                    out["solid_count"] = np.array([int(ref["solid"].sum())])
            print("  ", name, "N", n)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        manifest[name] = {"case": case.__dict__, "steps": steps_list, "kind": "sampled", "n_points": int(gx.size)}
        print("wrote", name)

    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
