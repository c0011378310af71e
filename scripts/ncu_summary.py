#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into one text table: duration, tensor pipe, DRAM throughput / bytes, warps
active, registers, per kernel launch.   python scripts/ncu_summary.py gpurun_out/a.ncu-rep [more ...] > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "dur_us"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor_inst_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("smsp__inst_executed.sum", "warp_inst")]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("== %s" % rep)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d.get("Kernel Name", "?")[:70]
            vals = []
            for k, short in KEYS:
                if k in d and d[k] != "":
                    v = d[k].replace(",", "")
                    try:
                        f = float(v)
                        if short in ("dram_rd", "dram_wr"):
                            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u.get(k, "byte"), 1)
                            vals.append("%s=%.1fMB" % (short, f * mult / 1e6))
                        elif short == "dur_us":
                            mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u.get(k, "ns"), 1e-3)
                            vals.append("%s=%.1f" % (short, f * mult))
                        else:
                            vals.append("%s=%.4g" % (short, f))
                    except ValueError:
                        pass
            print("%-72s %s" % (name, " ".join(vals)))


if __name__ == "__main__":
    main()
