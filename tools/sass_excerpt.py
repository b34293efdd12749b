"""SASS evidence of one kernel of the shipped library: instruction mix, the TMA bulk copies and their mbarriers,
packed-fp32 arithmetic, and the unrolled merge-walk loop.

    python tools/sass_excerpt.py <mangled-kernel-name> [lib.so] > profiles/<round>_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    kernel = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "1d-spectral-optimal-transport_b200", "_lib", "libsot_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, lib], capture_output=True, text=True, check=True).stdout
    ins = []
    for ln in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    demangled = subprocess.run(["c++filt", kernel], capture_output=True, text=True).stdout.strip()
    print(f"SASS of {demangled}\n(cuobjdump -sass of {os.path.relpath(lib, ROOT)}, sm_100a; {len(ins)} instructions in the kernel)\n")
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    print("Instruction mix of the whole kernel (static counts): " + ", ".join(f"{k}:{v}" for k, v in mix.most_common(24)) + "\n")
    print("TMA bulk copies and their mbarriers (cp.async.bulk -> UBLKCP, mbarrier -> SYNCS):")
    for a, t in ins:
        if "UBLKCP" in t or "SYNCS" in t:
            print(f"/*{a:04x}*/  {t} ;")
    print("\nPacked fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2), first occurrences:")
    shown = 0
    for a, t in ins:
        if re.search(r"\b(FFMA2|FADD2|FMUL2)\b", t) and shown < 10:
            print(f"/*{a:04x}*/  {t} ;")
            shown += 1
    # the merge-walk loop: the backward branch whose body holds the most LDS+STS pairs with FMNMX (4x unrolled)
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            body = [x for _, x in ins[j:i + 1]]
            score = sum("FMNMX" in x for x in body)
            if "LDS" in " ".join(body) and "STS" in " ".join(body) and 3 <= score <= 6 and (best is None or len(body) > best[2]):
                best = (j, i, len(body))
    if best:
        j, i, n = best
        print(f"\nMerge walk, gradient mode, uniform grid -- the whole 4x unrolled loop ({n} instructions = {n / 4:.1f} per merged slot):")
        for a, t in ins[j:i + 1]:
            print(f"/*{a:04x}*/  {t} ;")


if __name__ == "__main__":
    main()
