"""Host<->device copy rates of this box with pinned memory (what bounds bench.py's `e2e` line):
H2D alone, D2H alone, and both directions at once on two streams."""
import json
import time

import torch


def main(mb=512, reps=5):
    n = mb * (1 << 20) // 4
    h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    run(True, True)
    out = {"mb": mb, "h2d_gbs": mb / 1024 / run(True, False), "d2h_gbs": mb / 1024 / run(False, True)}
    t = run(True, True)
    out["both_each_gbs"] = mb / 1024 / t
    print(json.dumps(out))


if __name__ == "__main__":
    main()
