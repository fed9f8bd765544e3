#!/usr/bin/env python
"""Golden force history of the reference's README case (Re = 204.7: 2048 x 512, tau = 0.6,
inlet_velocity = 0.1333, 120 000 steps, output_frequency = 140).  TEST INFRASTRUCTURE.

Runs oracle/_ref/lbm_ref_fast -- the UNMODIFIED reference headers built with the reference's own
flags -- with OMP_NUM_THREADS=1 (the reference's boundary loops race at two corner
cells when threaded, so only the 1-thread run is reproducible; about 2 hours of CPU) and stores its
forces.csv as tests/golden/re200_forces_reference.csv.gz plus the Strouhal / C_D / C_L summary that
tools/strouhal.py (the headless scripts/lift.py) derives from it.

  python oracle/gen_golden_re200.py            # run the reference, then write the fixture
  python oracle/gen_golden_re200.py --from DIR # DIR already holds forces.csv of that run

tests/golden/re200_forces_strict_prefix.csv.gz is the head of forces.csv of the same case run with
oracle/_ref/lbm_ref_strict (OMP_NUM_THREADS=1; stopped after ~45 000 steps, complete rows kept).
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
CASE = ["--nx", "2048", "--ny", "512", "--uin", "0.1333", "--steps", "120000", "--of", "140"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from", dest="src", default=None)
    a = ap.parse_args()
    d = a.src or tempfile.mkdtemp(prefix="re200_")
    if not a.src:
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "lbm_ref_fast")] + CASE + ["--dump", d], env=env, check=True)
    text = open(os.path.join(d, "forces.csv")).read()
    rows = text.splitlines()
    assert rows[0].startswith("timestep,") and len(rows) == 1 + 858, len(rows)  # t = 0, 140, ..., 119980
    out = os.path.join(ROOT, "tests", "golden", "re200_forces_reference.csv.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(text.encode())
    # final fields of the same run (ref_harness --dump), sampled every 16th cell, for a field-level check
    import numpy as np

    if os.path.exists(os.path.join(d, "rho.bin")):
        samp = {k: np.fromfile(os.path.join(d, k + ".bin")).reshape(512, 2048)[::16, ::16].copy() for k in ("rho", "ux", "uy")}
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "re200_final_fields_sampled.npz"), stride=16, **samp)
    import strouhal

    tmp = os.path.join(d, "_forces_copy.csv")
    open(tmp, "w").write(text)
    res = strouhal.analyse(strouhal.load_forces(tmp), U=0.1333, D=2.0 * int(0.05 * 512))
    res["case"] = {"nx": 2048, "ny": 512, "tau": 0.6, "inlet_velocity": 0.1333, "steps": 120000, "output_frequency": 140,
                   "build": "oracle/_ref/lbm_ref_fast (-O3 -ffast-math -mavx2 -mfma -fopenmp), OMP_NUM_THREADS=1"}
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "re200_summary.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
