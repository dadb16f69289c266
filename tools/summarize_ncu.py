#!/usr/bin/env python
"""Key metrics of `ncu --set full` reports (one launch each) -> text for profiles/.

    python tools/summarize_ncu.py "title" report.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_top_kernels.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__cluster_dim_x",
]


def main():
    print(f"# {sys.argv[1]}")
    print("# ncu --set full --clock-control none, one launch each; cold-cache, serialised replay: durations are NOT bench values")
    for rep in sys.argv[2:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"\n== {rep.split('/')[-1]}")
            print(f"   kernel: {d.get('Kernel Name', '?')}")
            for m in METRICS:
                if m in d:
                    print(f"   {m:100s} {d[m]:>16s} {u[m]}")


if __name__ == "__main__":
    main()
