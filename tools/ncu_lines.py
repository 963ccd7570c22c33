"""Per source line executed warp instructions + stall samples from `ncu --page source --csv --print-source cuda,sass`
(first kernel of the report only)."""
import csv
import sys


def main():
    path, nq = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    rows = list(csv.reader(open(path)))
    cur_file, hdr, out, funcs = None, None, [], 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[2] == "-":  # a source line row (SASS rows carry an address)
            ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            try:
                out.append((cur_file, int(r[0]), r[1].strip(), int(r[ie]), int(r[smp])))
            except ValueError:
                pass
    # the report repeats files per kernel; keep the first half if duplicated
    seen, uniq = set(), []
    for o in out:
        k = (o[0], o[1])
        if k in seen:
            continue
        seen.add(k)
        uniq.append(o)
    tot = sum(o[3] for o in uniq)
    ts = sum(o[4] for o in uniq) or 1
    print(f"total {tot} ({tot / nq:.0f}/unit)")
    for f, ln, src, n, s in sorted(uniq, key=lambda o: -o[3])[:topn]:
        print(f"{n / nq:9.1f}/unit {100 * n / tot:5.1f}% | samples {100 * s / ts:5.1f}% | {f}:{ln}: {src[:110]}")


if __name__ == "__main__":
    main()
