"""Per-code-region view of an ncu source-page dump joined with nvdisasm line info.

    python tools/ncu_segments.py <src.csv> <all.sass> <section-index> <mangled-substring> <frames> <warps/frame> [chunk]
Prints, per chunk of SASS instructions: share of executed warp-instructions, instructions per thread per
frame, stall-sample share, shared-memory wavefronts per frame and the source-line span.
"""
import collections
import csv
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_by_line import line_table  # noqa: E402


def load(src_csv, section):
    rows = list(csv.reader(open(src_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    i = starts[section]
    hdr = rows[i + 1]
    body = []
    for r in rows[i + 2:]:
        if r and r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    return rows[i][1], hdr, body, len(starts)


def main():
    src_csv, sass, section, mangled, frames, wpf = sys.argv[1:7]
    chunk = int(sys.argv[7]) if len(sys.argv) > 7 else 96
    section, frames, wpf = int(section), int(frames), int(wpf)
    name, hdr, body, nsec = load(src_csv, section)
    ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    t = line_table(sass, mangled)
    print(name[:110], "| sections:", nsec, "| sass rows", len(body), "table", len(t))
    assert len(t) == len(body), "SASS table and profile disagree: rebuild the library the profile was taken with"
    tot = sum(int(r[ci] or 0) for r in body)
    tots = sum(int(r[cs] or 0) for r in body)
    per_thread = tot / frames / wpf
    print(f"warp-instr {tot}  per thread per frame {per_thread:.0f}")
    for a in range(0, len(body), chunk):
        b = min(a + chunk, len(body))
        inst = sum(int(r[ci] or 0) for r in body[a:b])
        samp = sum(int(r[cs] or 0) for r in body[a:b])
        wv = sum(int(r[cw] or 0) for r in body[a:b])
        if inst / tot < 0.003 and samp / max(tots, 1) < 0.003:
            continue
        lines = [t[k][1][1] for k in range(a, b) if t[k][1] and t[k][1][0] == "sot_kernels.cuh" and t[k][1][1] >= 330]
        ops = collections.Counter()
        for k in range(a, b):
            op = t[k][2].split()[0]
            if op.startswith("@"):
                op = t[k][2].split()[1]
            ops[op.split(".")[0]] += int(body[k][ci] or 0) / frames / wpf
        top = " ".join(f"{o}:{v:.0f}" for o, v in ops.most_common(7))
        print(f"{a:5d}-{b:5d} inst {100 * inst / tot:5.1f}% ({inst / frames / wpf:5.0f}/thr) samp {100 * samp / tots:5.1f}% "
              f"wave {wv / frames:6.1f}/frame lines {min(lines) if lines else '-'}..{max(lines) if lines else '-'} | {top}")


if __name__ == "__main__":
    main()
