"""Hottest CUDA source lines of one kernel in an ncu report (warp-stall samples per line).

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_hot_lines.py src.csv [table_index] [top_n]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    tables, hdr, cur, name = [], None, [], ""
    for r in rows:
        if r and r[0] == "Function Name":
            name = r[1]
        if r and r[0] == "Line No":
            if hdr is not None:
                tables.append((tname, hdr, cur))
            hdr, cur, tname = r, [], name
            continue
        if hdr and len(r) == len(hdr):
            cur.append(r)
    if hdr is not None:
        tables.append((tname, hdr, cur))
    tables = [t for t in tables if "Address" in t[1] and t[2] and t[2][0][t[1].index("Address")] == "-"]
    tname, hdr, out = tables[want]
    print("#", tname[:100], f"({len(tables)} source tables)")
    num = lambda s: int(s) if s.lstrip("-").isdigit() else 0
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_math", "stall_mio"]
    idx = [hdr.index(c) for c in cols]
    tot = sum(num(r[isamp]) for r in out) or 1
    print("# total samples", tot, " columns: samples share instr |", " ".join(c[6:] for c in cols))
    for r in sorted(out, key=lambda r: -num(r[isamp]))[:top]:
        print(f"{num(r[isamp]):6d} {100 * num(r[isamp]) / tot:5.1f}% {num(r[iex]):>9d} | "
              + " ".join(f"{num(r[i]):5d}" for i in idx) + f" | L{r[0]}: {r[1][:100]}")


if __name__ == "__main__":
    main()
