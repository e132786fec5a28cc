"""Key metrics + hottest source lines of an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                print(f"  {w:75s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        stalls = [(float(r[i] or 0), hdr[i]) for i in range(len(hdr)) if "issue_stalled" in hdr[i] and hdr[i].endswith("per_issue_active.ratio")]
        for v, name in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {name.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:8.2f} warps/issue")


def source(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if "Source" in r and ("Warp Stall Sampling (All Samples)" in r or "# Samples" in r or "Warp Stall Sampling (All Cycles)" in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if not hdr:
        print("no source page"); return
    col = next(c for c in ("Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)", "# Samples") if c in hdr)
    ci, si = hdr.index(col), hdr.index("Source")
    ex = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
    tot = sum(float(r[ci] or 0) for r in data) or 1
    print(f"  hottest SASS by {col} (total {tot:.0f}):")
    for r in sorted(data, key=lambda r: -float(r[ci] or 0))[:top]:
        print(f"   {100 * float(r[ci] or 0) / tot:5.1f}%  {r[si][:100]}" + (f"   [exec {r[ex]}]" if ex is not None else ""))


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
