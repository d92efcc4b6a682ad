#!/usr/bin/env python3
"""Counts the SASS opcodes inside the timed loop of every kernel of tools/imad_peak (cuobjdump -sass): the loop is the range
between the backward `BRA.U` and its target.  Output: one JSON line per mode — this is the evidence for `wide_per_iter_sass`."""
import collections, json, re, subprocess, sys, os
exe = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "imad_peak")
sass = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"^\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
    if m and cur:
        body[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
for f, ins in sorted(body.items()):
    loops = [(a, int(re.search(r"0x([0-9a-f]+)", rest).group(1), 16)) for a, op, rest in ins
             if op.startswith("BRA") and re.search(r"0x([0-9a-f]+)", rest) and int(re.search(r"0x([0-9a-f]+)", rest).group(1), 16) < a]
    a_end, a_start = loops[0]
    c = collections.Counter(op for a, op, _ in ins if a_start <= a <= a_end)
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    print(json.dumps({"kernel": f, "mode": int(re.search(r"ILi(\d+)E", f).group(1)), "imad_wide_per_iter": wide, "loop_opcodes": dict(c)}))
