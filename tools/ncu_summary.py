"""tools/ncu_summary.py <report.ncu-rep> [...] -- the handful of numbers the roofline claims rest on, as text (a
readable companion of the .ncu-rep files committed under profiles/)."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__cycles_elapsed.max",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        print(f"== {rep}: {rec.get('Kernel Name', '?')}  grid {rec.get('Grid Size', '?')} block {rec.get('Block Size', '?')}")
        for k in WANT:
            if k in rec:
                print(f"   {k:70s} {rec[k]} {units[hdr.index(k)]}")
        stalls = sorted(((float(v), h) for h, v in rec.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v), reverse=True)
        print("   top stall reasons (warps per issue-active cycle): " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in stalls[:6]))
