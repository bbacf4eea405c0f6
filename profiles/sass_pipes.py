#!/usr/bin/env python
"""Static pipe mix of the innermost (largest) loop of a kernel: ALU-pipe vs FMA-pipe vs other instruction counts.
On sm_100 the ALU pipe (LOP3, PRMT, SHF, VIADD*, VIMNMX*, VIADDMNMX*, ISETP, SEL, IADD3, LEA) issues one warp
instruction per two cycles per scheduler, the FMA pipes (IMAD*, FFMA...) likewise, so for an issue-bound loop
max(2*alu, 2*fma, total) approximates cycles per iteration.  Usage: sass_pipes.py <obj> <mangled-name substring>"""
import re
import subprocess
import sys
from collections import Counter

ALU = ("LOP3", "PRMT", "SHF", "VIADD", "VIMNMX", "VIADDMNMX", "ISETP", "SEL", "IADD3", "LEA", "IABS", "PLOP3", "FLO", "POPC", "BREV")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "IDP")


def main(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", out)[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)([^;]*);", f)
        loops = []
        for a, op, rest in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m and int(m.group(1), 16) < int(a, 16):
                    loops.append((int(m.group(1), 16), int(a, 16)))
        print(name[:110])
        for lo, hi in sorted(loops, key=lambda t: t[0]):
            body = [op.split(".")[0] for a, op, _ in ins if lo <= int(a, 16) <= hi]
            if len(body) < 200 or len(body) > 3000:
                continue
            c = Counter()
            for op in body:
                c["alu" if op.startswith(ALU) else "fma" if op.startswith(FMA) else "other"] += 1
            est = max(2 * c["alu"], 2 * c["fma"], len(body))
            print(f"  loop 0x{lo:x}-0x{hi:x}: {len(body):5d} instr  alu {c['alu']:4d}  fma {c['fma']:4d}  other {c['other']:4d}  "
                  f"=> >= {est} issue cycles/iter (static, both branches of any if counted)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
