#!/usr/bin/env python
"""Summarise an .ncu-rep (brought back in gpurun_out/) into profiles/<tag>.json + .md.
    python scripts/summarize_ncu.py gpurun_out/r01a_render.ncu-rep r01a_render_kernel [launches.csv]"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    # the L1 data pipe is what bounds a gather kernel: requests, tag sectors, data-stage wavefronts
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"report": os.path.basename(rep), "kernels": []}
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{vals[i]} {units[i]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len(STALL):-len("_per_issue_active.ratio")]] = float(vals[i].replace(",", ""))
                except ValueError:
                    pass
        d["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out["kernels"].append(d)
    os.makedirs("profiles", exist_ok=True)
    json.dump(out, open(f"profiles/{tag}.json", "w"), indent=1)
    with open(f"profiles/{tag}.md", "w") as f:
        f.write(f"# ncu summary `{tag}` (from {os.path.basename(rep)}; `ncu --set full --clock-control none`)\n\n")
        for d in out["kernels"]:
            f.write(f"## {d['kernel']}\n\n| metric | value |\n|---|---|\n")
            for k, v in d.items():
                if k not in ("kernel", "stall_cycles_per_issue"):
                    f.write(f"| {k} | {v} |\n")
            f.write("\nTop stall reasons (warp-cycles per issued instruction): " +
                    ", ".join(f"{k} {v:.2f}" for k, v in d["stall_cycles_per_issue"].items()) + "\n\n")
    print(open(f"profiles/{tag}.md").read())


if __name__ == "__main__":
    main()
