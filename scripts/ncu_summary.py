"""Summarise an .ncu-rep (ncu --set full) per kernel launch: duration, occupancy, issue / tensor-pipe / DRAM / L2
utilisation and the top warp-stall reasons.  usage: python scripts/ncu_summary.py file.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "dur_us": "gpu__time_duration.sum", "grid": "launch__grid_size", "block": "launch__block_size",
    "regs": "launch__registers_per_thread", "smem_dyn_kb": "launch__shared_mem_per_block_dynamic",
    "occ_limit_regs": "launch__occupancy_limit_registers", "occ_limit_smem": "launch__occupancy_limit_shared_mem",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_inst": "sm__inst_executed_pipe_tensor.sum",
    "inst_executed": "smsp__inst_executed.sum",
    "dram_read_mb": "dram__bytes_read.sum", "dram_write_mb": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]][:90]}
        for k, m in WANT.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                try:
                    v = float(r[idx[m]].replace(",", ""))
                    u = units[idx[m]]
                    if k in ("dram_read_mb", "dram_write_mb"):
                        v = v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                    if k == "dur_us":
                        v = v * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
                    d[k] = round(v, 3)
                except ValueError:
                    d[k] = r[idx[m]]
        st = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(r[i])))
                except ValueError:
                    pass
        st.sort(key=lambda x: -x[1])
        d["stalls_per_issue"] = {k: round(v, 2) for k, v in st[:7]}
        out.append(d)
    txt = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
