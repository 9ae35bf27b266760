"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py rep.ncu-rep > profiles/x.md"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{path}` (ncu --set full --clock-control none; cold-cache, serialised replay)\n")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')][:110]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        print()
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        ix = {k: i for i, k in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr) and (r[ix["# Samples"]] or "0").isdigit()]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        stall = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
        agg = sorted(((sum(int(r[ix[k]] or 0) for r in data), k) for k in stall), reverse=True)[:6]
        print("## warp-stall samples (all warps)\n")
        print(", ".join(f"{k}: {100 * v / tot:.1f}%" for v, k in agg), "\n")
        print("## hottest SASS instructions (by samples)\n\n| samples % | executed | SASS |\n|---|---|---|")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            print(f"| {100 * int(r[ix['# Samples']] or 0) / tot:.1f} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:90]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
