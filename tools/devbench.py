"""Development sweep (not the judged bench): MLUPS and algorithmic GB/s per kernel variant and size."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lbm_b200  # noqa: E402


def run(nx, ny, variant, steps=100, warm=10, flags=0, per_kernel=True):
    p = lbm_b200.SimulationParams(nx=nx, ny=ny, flags=flags, output_frequency=140)
    s = lbm_b200.Solver(p)
    s.set_kernel_variant(variant)
    s.initialise()
    s.step(warm)
    s.sync()
    ms, _, launches = s.time_steps(steps, False)
    out = {"nx": nx, "ny": ny, "variant": variant, "flags": flags, "steps": steps, "ms_per_step": ms / steps,
           "mlups": nx * ny * steps / ms / 1e3, "gbs_algo": nx * ny * 144 * steps / ms / 1e6, "launches": launches}
    if per_kernel:
        ms2, msb, _ = s.time_steps(20, True)
        out["bulk_ms"] = msb / 20
        out["bulk_gbs"] = nx * ny * 144 / (msb / 20) / 1e6
    ok, bad = s.check_stability()
    out["stable"] = ok
    s.close()
    return out


if __name__ == "__main__":
    sizes = [(2048, 512), (8192, 2048), (4096, 8192), (16384, 4096)]
    variants = [int(v) for v in os.environ.get("VARIANTS", "0,1").split(",")]
    for nx, ny in sizes:
        for v in variants:
            try:
                print(json.dumps(run(nx, ny, v)), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"nx": nx, "ny": ny, "variant": v, "error": str(e)}), flush=True)
