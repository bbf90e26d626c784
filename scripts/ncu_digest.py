#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> the handful of counters DESIGN.md / profiles/*.md quote, as a markdown table.
  python scripts/ncu_digest.py profiles/r02_ncu_pool_default_s8_raw.csv [more.csv ...]"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "live lanes per warp instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread instructions"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread instructions"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "of which bank conflicts"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (warp inst)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (warp inst)"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def digest(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    out = [f"### `{path.split('/')[-1]}` — {d['Kernel Name'][1].replace('void unnamed>::', '')}, block {d['Block Size'][1]}, grid {d['Grid Size'][1]}", "", "| counter | value |", "|---|---|"]
    for k, name in WANT:
        if k in d and d[k][1] != "":
            u, v = d[k]
            try:
                v = f"{float(v):,.2f}" if abs(float(v)) < 1e6 else f"{float(v):.4g}"
            except ValueError:
                pass
            out.append(f"| {name} (`{k}`) | {v} {u} |")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(d[h][1]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    out.append("| stalled warps per issued instruction, top 6 | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]) + " |")
    return "\n".join(out) + "\n", d


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(digest(p)[0])
