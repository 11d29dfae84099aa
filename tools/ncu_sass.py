#!/usr/bin/env python3
"""SASS-level view of an `ncu --page source --csv` export: how much of a kernel's executed instructions / stall samples sit in
its hot loop (instructions executed at least `thr` times), per opcode and per stall reason.
usage: ncu_sass.py <src.csv> [thr]"""
import collections
import csv
import re
import sys

csv.field_size_limit(10 ** 9)


def main(path, thr=20000):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    isrc, ismp, iex, ith = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        data.append((r[isrc], int(r[ismp] or 0), int(r[iex] or 0), int(r[ith] or 0), [int(r[i] or 0) for i, _ in stalls]))
    tot_i, tot_s = sum(d[2] for d in data) or 1, sum(d[1] for d in data) or 1
    print("%d SASS instructions in the kernel, %d executed warp-instructions, %d samples" % (len(data), tot_i, tot_s))
    for name, sel in (("hot (>= %d executions)" % thr, [d for d in data if d[2] >= thr]), ("cold", [d for d in data if 0 < d[2] < thr])):
        print("%-28s %6d SASS instrs  %5.1f%% of executed  %5.1f%% of samples  lanes %.1f" % (
            name, len(sel), 100 * sum(d[2] for d in sel) / tot_i, 100 * sum(d[1] for d in sel) / tot_s,
            sum(d[3] for d in sel) / max(1, sum(d[2] for d in sel))))
    hot = [d for d in data if d[2] >= thr]
    op_s, op_i = collections.Counter(), collections.Counter()
    for d in hot:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d[0])
        o = m.group(2).split(".")[0] if m else "?"
        op_s[o] += d[1]; op_i[o] += d[2]
    print("hot, by samples:", ", ".join("%s %.1f%%" % (k, 100 * v / tot_s) for k, v in op_s.most_common(12)))
    print("hot, by executed:", ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in op_i.most_common(12)))
    st = collections.Counter()
    for d in hot:
        for (i, h), v in zip(stalls, d[4]):
            st[h] += v
    tot = sum(st.values()) or 1
    print("hot, stall reasons:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / tot) for k, v in st.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20000)
