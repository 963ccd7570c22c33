"""From `ncu --page source --csv --print-source cuda,sass`: for the given SASS opcodes, the source lines that execute them most.
usage: python tools/ncu_ops_by_line.py src.csv units OP1,OP2,... [topn]"""
import collections
import csv
import re
import sys


def main():
    path, nq, ops = sys.argv[1], float(sys.argv[2]), set(sys.argv[3].split(","))
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    cur_file, hdr, cur_line, cur_src = None, None, None, ""
    per = collections.defaultdict(collections.Counter)
    seen_files = set()
    skip = False
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            skip = r[1] in seen_files  # the report repeats the files per kernel
            seen_files.add(r[1])
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            continue
        if skip or not hdr or len(r) != len(hdr):
            continue
        if r[2] == "-":
            cur_line, cur_src = r[0], r[1].strip()
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
        if not m:
            continue
        op = m.group(2)
        if op in ops:
            try:
                per[op][(cur_file, cur_line, cur_src)] += int(r[ie])
            except ValueError:
                pass
    for op in ops:
        tot = sum(per[op].values())
        print(f"== {op}: {tot / nq:.0f}/unit")
        for (f, ln, src), c in per[op].most_common(topn):
            print(f"   {c / nq:8.1f}  {f}:{ln}: {src[:100]}")


if __name__ == "__main__":
    main()
