#!/usr/bin/env python
"""Where each warp role of a kernel spends its time: stall samples of an ncu source page, summed between synchronisation
points (mbarrier waits, named barriers, tcgen05 MMA / commit, tensor-memory loads / stores, global stores).

    ncu -i prof.ncu-rep --page source --csv > src.csv      # the report needs --set full --import-source on
    python scripts/ncu_segments.py src.csv [min-share-percent]

A role's share of all samples is its share of the resident warps (e.g. 3 prep warps of 20 = 15 %): a wait segment that
holds a third of the epilogue warps' samples is a third of their time (profiles/README.md)."""
import csv
import sys

MARKERS = ("BAR.SYNC", "BAR.ARV", "TRYWAIT", "UTCHMMA", "UTCBAR", "ARRIVE", "LDTM", "STTM", "EXIT")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    hdr, data = rows[1], rows[2:]
    i_samp, i_exec = hdr.index("# Samples"), hdr.index("Instructions Executed")
    total = sum(int(r[i_samp] or 0) for r in data)
    print("kernel:", rows[0][1][:100])
    print("total samples", total)
    acc = 0
    for r in data:
        acc += int(r[i_samp] or 0)
        if any(m in r[1] for m in MARKERS) or ("STG" in r[1] and acc > 100):
            if acc > total * floor / 100 or "BAR" in r[1]:
                print(f"{r[0][-5:]} {acc:8d} {100 * acc / total:5.1f}%  executed {r[i_exec]:>10s}  {r[1].strip()[:72]}")
            acc = 0


if __name__ == "__main__":
    main()
