"""Summarise an `ncu --page source --csv` dump: executed warp instructions and stall samples per SASS opcode."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data, seen = None, [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            seen += 1
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if seen == 1 and hdr and len(r) == len(hdr):
            data.append(r)
    return hdr, data


def main():
    path, nq = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr, data = load(path)
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    tot = sum(int(r[ia]) for r in data)
    print("total warp instructions", tot, "per unit", tot / nq, "| SASS lines", len(data))
    op, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[ia])
        samp[o] += int(r[isamp])
    ts = sum(samp.values()) or 1
    for o, c in op.most_common(34):
        print(f"{o:10s} {c / nq:10.1f}/unit {100 * c / tot:5.1f}%   stall samples {100 * samp[o] / ts:5.1f}%")


if __name__ == "__main__":
    main()
