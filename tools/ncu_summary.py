"""Text summary of an .ncu-rep (the metrics DESIGN.md / bench.py cite), one block per profiled launch.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_full.txt"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none  (%s)" % rep.split("/")[-1])
    tot_r = tot_w = n = 0
    for r in rows[2:]:
        print("launch %s  %s" % (r[idx["ID"]], r[idx["Kernel Name"]][:150]))
        for w in WANT:
            if w in idx:
                print("    %-78s %s %s" % (w, r[idx[w]], units[idx[w]]))
        def byt(name):
            v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tot_r += byt("dram__bytes_read.sum"); tot_w += byt("dram__bytes_write.sum"); n += 1
    print("# mean DRAM bytes per launch: read %.3f MB, write %.3f MB, total %.3f MB over %d launches" % (tot_r / n / 1e6, tot_w / n / 1e6, (tot_r + tot_w) / n / 1e6, n))
if __name__ == "__main__":
    main()
