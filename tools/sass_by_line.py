"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm -g` line info and
aggregate executed warp-instructions / stall samples per source line (and per stage of the kernel).

    python tools/sass_by_line.py <ncu_sass.csv> <nvdisasm_all.sass> <kernel-substring> [--top 40]
"""
import csv
import re
import sys
from collections import defaultdict


def line_table(sass_path, kernel):
    table, cur, active = [], None, False
    for ln in open(sass_path, errors="replace"):
        if ln.startswith(".text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return table


def main():
    ncu_csv, sass_path, kernel = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    rows = list(csv.reader(open(ncu_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sel = None
    for k, i in enumerate(starts):
        name = rows[i][1].replace("(int)", "").replace("(bool)", "").replace(" ", "")
        if kernel in name:
            sel = (i, starts[k + 1] if k + 1 < len(starts) else len(rows))
            print("kernel:", name)
    assert sel, "kernel not found in ncu dump: " + str([rows[i][1] for i in starts])
    hdr = rows[sel[0] + 1]
    c_inst, c_samp, c_src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    c_wave = hdr.index("L1 Wavefronts Shared")
    body = rows[sel[0] + 2:sel[1]]
    mangled = sys.argv[sys.argv.index("--mangled") + 1] if "--mangled" in sys.argv else None
    table = line_table(sass_path, mangled or kernel)
    assert len(table) >= len(body) - 2, (len(table), len(body))
    per_line = defaultdict(lambda: [0, 0, 0])
    total = [0, 0, 0]
    for i, r in enumerate(body):
        if i >= len(table):
            break
        inst, samp, wave = int(r[c_inst] or 0), int(r[c_samp] or 0), int(r[c_wave] or 0)
        key = table[i][1]
        for acc in (per_line[key], total):
            acc[0] += inst
            acc[1] += samp
            acc[2] += wave
    print(f"total warp-instr {total[0]}, samples {total[1]}, smem wavefronts {total[2]}")
    print(f"{'file:line':28s} {'inst%':>7s} {'samp%':>7s} {'wave%':>7s}")
    for key, (inst, samp, wave) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{key[0] + ':' + str(key[1]):28s} {100 * inst / total[0]:7.2f} {100 * samp / max(total[1], 1):7.2f} "
              f"{100 * wave / max(total[2], 1):7.2f}")
    if "--ranges" in sys.argv:
        spec = sys.argv[sys.argv.index("--ranges") + 1]  # "name:lo-hi,name:lo-hi" on sot_kernels.cuh
        print("---- stages")
        for item in spec.split(","):
            nm, rng = item.split(":")
            lo, hi = (int(t) for t in rng.split("-"))
            agg = [0, 0, 0]
            for (f, ln), v in per_line.items():
                if f == "sot_kernels.cuh" and lo <= ln <= hi:
                    for k in range(3):
                        agg[k] += v[k]
            print(f"{nm:24s} inst {100 * agg[0] / total[0]:6.2f}%  samples {100 * agg[1] / max(total[1], 1):6.2f}%  "
                  f"wavefronts {100 * agg[2] / max(total[2], 1):6.2f}%")


if __name__ == "__main__":
    main()
