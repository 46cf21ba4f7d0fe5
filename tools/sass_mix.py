#!/usr/bin/env python
"""Static evidence for the access pattern of every kernel in libsfron_b200.so: per-kernel counts of 128-bit vs
narrower global loads / stores, shared / global atomics, warp shuffles, barriers and tensor-core instructions in the
sm_100a SASS (`cuobjdump -sass`).  Runs without a GPU:  python tools/sass_mix.py > profiles/r1_sass_instruction_mix.txt
(narrow loads are the 4-byte packed mask words, scalars and the ragged tails; there is no MMA on this path)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unified-unlearning-w-remain-geometry_b200", "libsfron_b200.so")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        mangled = part.split("\n", 1)[0].strip()
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("sfr::(anonymous namespace)::", "")).replace("void ", "")
        ops = collections.Counter(m.group(1) for m in re.finditer(
            r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part, flags=re.M))
        total = lambda pred: sum(v for k, v in ops.items() if pred(k))
        rows.append((name,
                     total(lambda k: k.startswith("LDG") and ".128" in k), total(lambda k: k.startswith("LDG") and ".128" not in k),
                     total(lambda k: k.startswith("STG") and ".128" in k), total(lambda k: k.startswith("STG") and ".128" not in k),
                     total(lambda k: k.startswith("ATOMS")), total(lambda k: k.startswith(("ATOMG", "RED"))),
                     total(lambda k: k.startswith("SHFL")), total(lambda k: k.startswith("BAR")),
                     total(lambda k: "MMA" in k or "TCGEN" in k or k.startswith("UTC"))))
    print(f"{'kernel':46s} LDG.128 LDG.other STG.128 STG.other ATOMS ATOMG/RED SHFL BAR MMA/tcgen05")
    for r in sorted(rows):
        print(f"{r[0][:46]:46s} {r[1]:7d} {r[2]:9d} {r[3]:7d} {r[4]:9d} {r[5]:5d} {r[6]:9d} {r[7]:4d} {r[8]:3d} {r[9]:11d}")


if __name__ == "__main__":
    main()
