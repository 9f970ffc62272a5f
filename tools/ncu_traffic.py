"""DRAM traffic per launch of the roofline kernels, from `ncu --page raw --csv` of
`tools/profile_targets.py gemm 1` (launch order = bench.ffn_instep_launches order), written to
profiles/r2_roofline_traffic.json, which bench.py reads for `roofline.traffic` (no literals).

    ncu -i gpurun_out/r2_gemm.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv profiles/r2_roofline_traffic.json [launches skipped by ncu -s]
"""
import csv
import json
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    names = ["ffn1_fwd", "ffn2_fwd", "ffn_dgrad"]
    out = {}
    # launches skipped by `ncu -s N` before the first captured one (the three launches cycle)
    k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    k = 0
    for r in rows[2:]:
        if "gemm_tc_kernel" not in r[col["Kernel Name"]]:
            continue
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        if k < len(names):
            out[names[(k + k0) % len(names)]] = rd + wr
        k += 1
    path = sys.argv[2]
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    cur["gemm_tc"] = out
    cur["source"] = ("ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 4 of `python "
                     "tools/profile_targets.py gemm 2` (launch order ffn1_fwd, ffn2_fwd, ffn_dgrad per round; "
                     "the first captured launch is the dgrad of round 1); summary: profiles/r2_gemm_tc_instep.txt")
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    print(out)


if __name__ == "__main__":
    main()
