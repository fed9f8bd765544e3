"""Static instruction mix of the loops of one kernel in a cuobjdump -sass listing (stdin or file):
python tools/sass_loops.py listing.sass   ->   every backward branch with the opcode histogram of its body."""
import collections
import re
import sys

lines = open(sys.argv[1]).read().splitlines() if len(sys.argv) > 1 else sys.stdin.read().splitlines()
ins = []
for l in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s*(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), (m.group(4) or ""), m.group(5)))
addr_index = {a: k for k, (a, _, _, _) in enumerate(ins)}
DP = {"DADD", "DMUL", "DFMA", "DSETP", "DMNMX"}
for k, (a, op, mod, rest) in enumerate(ins):
    if op != "BRA":
        continue
    m = re.search(r"0x([0-9a-f]+)", rest)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t >= a or t not in addr_index:
        continue
    body = ins[addr_index[t]:k + 1]
    h = collections.Counter(o for _, o, _, _ in body)
    dp = sum(v for o, v in h.items() if o in DP)
    print("loop 0x%x..0x%x: %d instr, DP %d, LDG %d, STG %d, LDS %d, STS %d, BAR %d" % (
        t, a, len(body), dp, h["LDG"], h["STG"], h["LDS"], h["STS"], h["BAR"]))
    print("   ", ", ".join("%s %d" % kv for kv in h.most_common(14)))
