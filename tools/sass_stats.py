#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in the built library:
    python tools/sass_stats.py [substring of mangled name] [--dump]"""
import collections, re, subprocess, sys
name = sys.argv[1] if len(sys.argv) > 1 else "scan_kernelILi2ELb1"
out = subprocess.run(["cuobjdump", "-sass", "pgrc_b200/libpgrc_gpu.so"], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    fn = b.split("\n", 1)[0].strip()
    if name not in fn: continue
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", b, flags=re.M)
    ops = collections.Counter()
    for i in ins:
        t = i.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    print(fn, len(ins), "instructions")
    print(ops.most_common(40))
    if "--dump" in sys.argv:
        print("\n".join(ins))
