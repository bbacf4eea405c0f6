#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS, split into the hot loop (between the last backward branch
target and the branch) and the rest.  Usage: sass_hist.py <object or .so> <substring of mangled name>"""
import re
import subprocess
import sys
from collections import Counter


def main(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)([^;]*);", f)
        addr = [int(a, 16) for a, _, _ in ins]
        # find backward branches
        loops = []
        for a, op, rest in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m and int(m.group(1), 16) < int(a, 16):
                    loops.append((int(m.group(1), 16), int(a, 16)))
        print(name[:120])
        print("  total instructions:", len(ins))
        for lo, hi in sorted(loops, key=lambda t: t[0] - t[1])[:3]:
            body = [op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("VI", "PRMT")) else "")
                    for a, op, _ in ins if lo <= int(a, 16) <= hi]
            print(f"  loop 0x{lo:x}-0x{hi:x}: {len(body)} instructions")
            for op, c in Counter(body).most_common(14):
                print(f"      {c:6d} {op}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
