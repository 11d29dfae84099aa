#!/usr/bin/env python3
"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep (ncu --import-source on, -lineinfo)."""
import csv
import subprocess
import sys


def main(rep, top=40, per=1):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    out = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        d = dict(zip(range(len(hdr)), r))
        ex = int(r[hdr.index("Instructions Executed")] or 0)
        th = int(r[hdr.index("Thread Instructions Executed")] or 0)
        sm = int(r[hdr.index("# Samples")] or 0)
        out.append((ex, th, sm, r[0], r[1]))
    tot_ex = sum(o[0] for o in out) or 1
    tot_sm = sum(o[2] for o in out) or 1
    print("total warp-instructions %d (%.0f per unit), samples %d" % (tot_ex, tot_ex / per, tot_sm))
    for ex, th, sm, ln, src in sorted(out, key=lambda o: -o[2])[:top]:
        print("%5.1f%% smp %5.1f%% ins  %8.1f/unit  lanes %4.1f  L%-4s %s" % (100 * sm / tot_sm, 100 * ex / tot_ex, ex / per, th / max(ex, 1), ln, src.strip()[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, float(sys.argv[3]) if len(sys.argv) > 3 else 1)
