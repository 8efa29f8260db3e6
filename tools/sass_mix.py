#!/usr/bin/env python
"""Instruction mix of one kernel from `cuobjdump -sass` (here, no GPU needed):

    python tools/sass_mix.py <object or .so> <substring of the mangled kernel name> [--top N]

Prints the number of SASS instructions per opcode of the FIRST function whose name contains the
substring -- the static count, which for the fully unrolled filter kernels is the per-thread
dynamic count of the straight-line part."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", text)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f):
            ops[m.group(1)] += 1
        total = sum(ops.values())
        print(f"{name}: {total} instructions")
        for op, n in ops.most_common(top):
            print(f"  {op:12s} {n:6d}  {100.0 * n / total:5.1f} %")
        return
    print("no function matches", pat)


if __name__ == "__main__":
    main()
