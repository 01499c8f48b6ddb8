"""Print the SASS of one kernel (substring match on the mangled name) and an opcode histogram.

    python tools/sass_of.py irs_mpc_b200/libirs_mpc_b200.so QuadrotorIfEELi4 [--hist]
"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat in name:
        if "--hist" in sys.argv:
            ops = collections.Counter()
            for line in b.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
                if m:
                    ops[m.group(1)] += 1
            print(name, sum(ops.values()))
            for k, v in ops.most_common(40):
                print("  %6d %s" % (v, k))
        else:
            print("Function : " + b)
