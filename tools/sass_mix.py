#!/usr/bin/env python
"""Static evidence for the access pattern of every kernel in libsfron_b200.so: per-kernel counts of 128-bit vs
narrower global loads / stores, shared / global atomics, warp shuffles, barriers and tensor-core instructions in the
sm_100a SASS (`cuobjdump -sass`).  Runs without a GPU:  python tools/sass_mix.py > profiles/r2_sass_instruction_mix.txt
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
                     total(lambda k: "MMA" in k or "TCGEN" in k or k.startswith("UTC")),
                     total(lambda k: k.startswith("UBLKCP")), total(lambda k: k.startswith("SYNCS")),
                     total(lambda k: k.startswith("LDGMC")), total(lambda k: ".STRONG.SYS" in k)))
    # kernels with many template instantiations (optimizer x EMA x dtype x source) are summed per kernel name
    merged = collections.OrderedDict()
    for r in sorted(rows):
        key = re.sub(r"<.*", "", r[0])
        acc = merged.setdefault(key, [0] * (len(r) - 1) + [0])
        for i, v in enumerate(r[1:]):
            acc[i] += v
        acc[-1] += 1
    print(f"{'kernel (instantiations summed)':40s} inst LDG.128 LDG.other STG.128 STG.other ATOMS ATOMG/RED SHFL BAR MMA "
          f"UBLKCP(TMA) SYNCS(mbarrier) LDGMC(multimem) .STRONG.SYS")
    for k, v in merged.items():
        print(f"{k[:40]:40s} {v[-1]:4d} {v[0]:7d} {v[1]:9d} {v[2]:7d} {v[3]:9d} {v[4]:5d} {v[5]:9d} {v[6]:4d} {v[7]:3d} {v[8]:3d} "
              f"{v[9]:11d} {v[10]:15d} {v[11]:15d} {v[12]:11d}")


if __name__ == "__main__":
    main()
