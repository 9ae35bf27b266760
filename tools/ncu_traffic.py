"""Write profiles/r02_traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch of the
dominant kernels (bench.py reads it for roofline.traffic).  python tools/ncu_traffic.py prefill.ncu-rep decode.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out = []
    for x in r[2:]:
        d = dict(zip(hdr, x))
        u = dict(zip(hdr, units))

        def val(k):
            v = float(d[k].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u[k], 1.0)
        out.append({"kernel": d["Kernel Name"], "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                    "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u["gpu__time_duration.sum"], 1.0),
                    "grid": d.get("launch__grid_size")})
    return out


def main(prefill_rep, decode_rep):
    out = {}
    if prefill_rep and os.path.exists(prefill_rep):
        rs = [r for r in rows(prefill_rep) if "gemm_tt" in r["kernel"]]
        if rs:
            r = rs[0]
            out["prefill_gemm"] = {"launch": "gemm_tt_kernel M=16384 N=4096 K=4096 (first captured launch)",
                                   "bytes_per_launch": r["dram_read"] + r["dram_write"], "dram_read": r["dram_read"], "dram_write": r["dram_write"],
                                   "algorithmic_bytes": 2 * 16384 * 4096 * 2 + 0.53 * 4096 * 4096,
                                   "capture": os.path.basename(prefill_rep), "duration_us_under_ncu": r["duration_us"]}
    if decode_rep and os.path.exists(decode_rep):
        rs = [r for r in rows(decode_rep) if "decode_" in r["kernel"]]
        if rs:
            tot = sum(r["dram_read"] + r["dram_write"] for r in rs)
            out["decode"] = {"launch": f"decode kernel ({rs[0]['kernel'][:40]}), average over the {len(rs)} captured launches (one decoder layer: q,k,v,o,gate,up,down), batch 8",
                             "bytes_per_launch": tot / len(rs), "per_launch": [r["dram_read"] + r["dram_write"] for r in rs],
                             "capture": os.path.basename(decode_rep)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, sys.argv[2] if len(sys.argv) > 2 else None)
