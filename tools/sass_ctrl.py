"""Decode the scheduling control bits (stall count, yield, scoreboard set/wait) of a kernel's SASS.
usage: python tools/sass_ctrl.py <lib.so> <function substring> [start_pattern] [lines]"""
import re, subprocess, sys
lib, fn = sys.argv[1], sys.argv[2]
pat = sys.argv[3] if len(sys.argv) > 3 else None
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 200
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on = False; rows = []; i = 0
while i < len(out):
    l = out[i]
    if "Function :" in l: on = fn in l
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if on and m and i + 1 < len(out):
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = hi >> 41          # bits 105.. of the 128-bit word
            stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wbar = (ctrl >> 5) & 7; rbar = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3f
            rows.append((m.group(1), stall, yld, wbar, rbar, wait, m.group(2).strip()))
            i += 1
    i += 1
start = 0
if pat:
    for k, r in enumerate(rows):
        if pat in r[6]: start = max(0, k - 5); break
tot = 0
for r in rows[start:start + nl]:
    tot += r[1]
    print(f"{r[0]} st={r[1]:2d} y={r[2]} w={r[3] if r[3]!=7 else '-'} r={r[4] if r[4]!=7 else '-'} wait={r[5]:06b} {r[6][:90]}")
print("sum of stall counts:", tot)
