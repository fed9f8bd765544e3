"""Dynamic instruction mix of the captured bulk kernel, from the SASS page of an ncu report:
  python tools/ncu_sass_mix.py gpurun_out/r02_bulk.ncu-rep <cell updates per launch> > profiles/r02_bulk_sass_mix.md
(executed warp instructions per opcode, per cell update, with the share of the stall samples each opcode drew)."""
import collections
import csv
import subprocess
import sys

rep, updates = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
name = next(csv.reader([lines[idx[0]]]))[1]
rows = list(csv.DictReader(lines[idx[0] + 1: idx[1] if len(idx) > 1 else None]))
count, samples = collections.Counter(), collections.Counter()
for r in rows:
    op = r["Source"].split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    count[o] += int(r["Instructions Executed"])
    samples[o] += int(r["# Samples"])
tot, stot = sum(count.values()), sum(samples.values())
per = updates / 32.0  # warp-level cell updates
DP = ("DADD", "DMUL", "DFMA", "DSETP")
MEM = ("LDG", "STG", "LDS", "STS", "CCTL", "UBLKPF")
print("# Executed instruction mix of `%s` (ncu SASS page, first captured launch)\n" % name)
print("%d warp instructions for %.0f cell updates = **%.1f instructions per cell update** (per thread); "
      "double-precision arithmetic %.1f, memory %.1f, everything else (addresses, ring, predicates, branches) %.1f.\n" % (
          tot, updates, tot / per, sum(count[o] for o in DP) / per, sum(count[o] for o in MEM) / per,
          (tot - sum(count[o] for o in DP + MEM)) / per))
print("| opcode | per cell update | share of instructions | share of stall samples |")
print("|---|---:|---:|---:|")
for o, v in count.most_common(28):
    print("| %s | %.2f | %.1f %% | %.1f %% |" % (o, v / per, 100.0 * v / tot, 100.0 * samples[o] / max(stot, 1)))

hot = sorted(rows, key=lambda r: -int(r["# Samples"]))[:40]
print("\n## The 40 SASS instructions that drew the most stall samples (of %d)\n" % stot)
print("| address | instruction | executed (warps) | samples |")
print("|---|---|---:|---:|")
for r in hot:
    print("| ...%s | `%s` | %s | %s |" % (r["Address"][-5:], " ".join(r["Source"].split()), r["Instructions Executed"], r["# Samples"]))
