#!/usr/bin/env python3
"""Summarise an ncu report (CPU side, no GPU needed): python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1_ld_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1_st_sectors"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "l2_write_sectors"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s` (`ncu --set full --clock-control none`)\n" % path.split("/")[-1])
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        print("## %s  (launch id %s)\n" % (name, r[ix["ID"]]))
        print("| metric | value | unit |\n|---|---|---|")
        for m, label in WANT:
            if m in ix:
                print("| %s (`%s`) | %s | %s |" % (label, m, r[ix[m]], units[ix[m]]))
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i] or 0)) for i, h in enumerate(hdr)
              if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued")]
        tot = sum(v for _, v in st) or 1.0
        top = sorted(st, key=lambda x: -x[1])[:5]
        print("\nstall samples: " + ", ".join("%s %.0f%%" % (h, 100 * v / tot) for h, v in top) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
