"""Sweep of the temporally blocked kernel's launch shape on one GPU (run under gpurun):
python tools/tb_sweep.py [workload]  ->  one line per (variant, depth, block rows, chunk width)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(workload, variant, depth, steps):
    import bench
    import lbm_b200

    cfg = bench.workload(workload, 1)
    p = lbm_b200.SimulationParams(nx=cfg["nx"], ny=cfg["ny"], tau=cfg["tau"], inlet_velocity=cfg["inlet_velocity"],
                                  output_frequency=cfg["output_frequency"], flags=cfg["flags"])
    s = lbm_b200.Solver(p)
    s.set_kernel_variant(variant)
    s.set_pass_depth(depth)
    s.initialise()
    s.step(13)
    s.sync()
    best = None
    for _ in range(3):
        ms, _, launches = s.time_steps(steps, 0)
        best = ms if best is None else min(best, ms)
    mlups = cfg["nx"] * cfg["ny"] * steps / (best * 1e-3) / 1e6
    print(json.dumps({"workload": workload, "variant": variant, "depth": depth, "B": os.environ.get("LBM_B200_TB_B", "256"),
                      "xc": os.environ.get("LBM_B200_TB_XC", "auto"), "ms_per_step": best / steps, "mlups": round(mlups),
                      "launches": launches}), flush=True)
    s.close()


if __name__ == "__main__":
    if len(sys.argv) > 2:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    wl = sys.argv[1] if len(sys.argv) > 1 else "slab"
    combos = [(1, 1, "256", "0"), (2, 1, "256", "0")]
    for depth in (2, 3):
        for B in ("256", "128"):
            if depth == 3 and B == "128":
                continue
            for xc in ("0", "32", "64", "128", "256", "512"):
                combos.append((2, depth, B, xc))
    for variant, depth, B, xc in combos:
        env = dict(os.environ, LBM_B200_TB_B=B)
        if xc != "0":
            env["LBM_B200_TB_XC"] = xc
        else:
            env.pop("LBM_B200_TB_XC", None)
        subprocess.run([sys.executable, os.path.abspath(__file__), wl, str(variant), str(depth), "120"], env=env)
