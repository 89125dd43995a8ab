"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into profiles/<name>.summary.txt:
key raw metrics of the captured kernel + the top stalled SASS instructions of the source page."""
import csv, io, os, subprocess, sys

rep, name = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
out = []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum ", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_sample_count"]
out.append(f"# ncu summary: {name}   ({note})")
out.append(f"# source report: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on; not a timing run)")
for h, u, v in zip(hdr, units, vals):
    if any((h + " ").startswith(k) or h == k.strip() for k in KEYS):
        out.append(f"{h} [{u}] = {v}")
out.append("")
out.append("## warp stall samples (smsp__pcsamp_warps_issue_stalled_*)")
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
        out.append(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    def g(r, k):
        try: return int(r[ix[k]])
        except Exception: return 0
    data = rows[2:]
    out.append("")
    out.append("## top 25 SASS instructions by stall samples: addr samples long_sb wait math short_sb barrier | instruction")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:25]:
        out.append(f"{r[ix['Address']][-5:]} {g(r, '# Samples'):8d} {g(r, 'stall_long_sb'):8d} {g(r, 'stall_wait'):8d} {g(r, 'stall_math'):8d} "
                   f"{g(r, 'stall_short_sb'):8d} {g(r, 'stall_barrier'):8d} | {r[ix['Source']][:70]}")
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{name}.summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
