"""Per-code-region view of an `ncu --page source --csv` dump joined with `nvdisasm -g` line info: executed
instructions per thread and frame, shared-memory wavefronts per frame, and stall samples by reason.

    python tools/ncu_stalls.py <src.csv> <all.sass> <mangled-substring> <frames> <warps/frame> [lines-per-region]
Regions are ranges of source lines of sot_kernels.cuh (inlined helpers are attributed to the caller's region by
SASS position)."""
import collections
import csv
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_by_line import line_table  # noqa: E402

REASONS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_long_sb", "stall_math", "stall_mio",
           "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_sleep",
           "stall_membar", "stall_lg"]


def main():
    src_csv, sass, mangled, frames, wpf = sys.argv[1:6]
    chunk = int(sys.argv[6]) if len(sys.argv) > 6 else 64
    frames, wpf = int(frames), int(wpf)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    body = [r for r in rows[2:] if r and r[0].startswith("0x")]
    t = line_table(sass, mangled)
    assert len(t) == len(body), (len(t), len(body))
    ci, cw = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
    cr = {k: hdr.index(k) for k in REASONS}
    tot = sum(int(r[ci] or 0) for r in body)
    tots = sum(int(r[cr[k]] or 0) for r in body for k in REASONS)
    print(f"warp-instr {tot}  per thread per frame {tot / frames / wpf:.0f}; stall samples {tots}")
    by = collections.Counter()
    for r in body:
        for k in REASONS:
            by[k] += int(r[cr[k]] or 0)
    print("all:", " ".join(f"{k[6:]}:{100 * v / tots:.1f}%" for k, v in by.most_common()))
    for a in range(0, len(body), chunk):
        b = min(a + chunk, len(body))
        inst = sum(int(r[ci] or 0) for r in body[a:b])
        wv = sum(int(r[cw] or 0) for r in body[a:b])
        st = collections.Counter()
        for r in body[a:b]:
            for k in REASONS:
                st[k] += int(r[cr[k]] or 0)
        samp = sum(st.values())
        if inst / tot < 0.003 and samp / tots < 0.003:
            continue
        lines = [t[k][1][1] for k in range(a, b) if t[k][1] and t[k][1][0] == "sot_kernels.cuh"]
        ops = collections.Counter()
        for k in range(a, b):
            parts = t[k][2].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += int(body[k][ci] or 0) / frames / wpf
        top = " ".join(f"{o}:{v:.0f}" for o, v in ops.most_common(6))
        sts = " ".join(f"{k[6:]}:{100 * v / tots:.1f}" for k, v in st.most_common(4))
        print(f"{a:5d}-{b:5d} inst {inst / frames / wpf:5.0f}/thr samp {100 * samp / tots:5.1f}% wave {wv / frames:6.1f} "
              f"L{min(lines) if lines else 0}-{max(lines) if lines else 0} | {top} | {sts}")


if __name__ == "__main__":
    main()
