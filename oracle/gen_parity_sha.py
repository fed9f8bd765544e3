"""Golden SHA-256 of the CPU oracle's result for bench.py's in-run parity check (tests/golden/bench_parity_sha.json).
TEST INFRASTRUCTURE.  Run here (CPU): python oracle/gen_parity_sha.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (parity_case / parity_state / parity_sha: the definitions the GPU run uses)
from oracle import oracle as O  # noqa: E402


def main():
    out = {}
    for n in (1, 2, 4, 8):
        c = bench.parity_case(n)
        case = O.Case(**c)
        o = O.Oracle(case)
        o.f_current[...] = bench.parity_state(c["nx"], c["ny"])
        rows, bad = o.run(bench.PARITY_STEPS)
        assert bad == -1
        out[str(n)] = {"sha256": bench.parity_sha(o.f_next[1:-1, 1:-1], rows), "case": c, "steps": bench.PARITY_STEPS,
                       "forces_rows": len(rows)}
    path = os.path.join(ROOT, "tests", "golden", "bench_parity_sha.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(open(path).read())


if __name__ == "__main__":
    main()
