"""Summarise an .ncu-rep (one kernel) into a small text file: python profiles/summarize_ncu.py rep out.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
lines = []
for kidx in range(2, len(rows)):
    lines.append(f"== kernel launch row {kidx - 2}: {rows[kidx][rows[0].index('Kernel Name')] if 'Kernel Name' in rows[0] else ''}")
    for h, u, v in zip(rows[0], rows[1], rows[kidx]):
        if any(h == k or h.startswith(k + " ") or (k in h and k.endswith("realtime")) for k in keys):
            lines.append(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(srows) if "# Samples" in r]
if hi:
    hdr = srows[hi[0]]; data = [r for r in srows[hi[0] + 1:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iS]) for r in data) or 1
    texec = sum(int(r[iEx]) for r in data)
    lines.append(f"== warp-stall sampling: {tot} samples, {texec} warp-instructions executed")
    st_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[c]: sum(int(r[c]) for r in data) for c in st_cols}
    lines.append("stall reasons (all warps): " + ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    lines.append("top SASS lines by samples:")
    for r in sorted(data, key=lambda r: -int(r[iS]))[:14]:
        main = sorted(((hdr[c], int(r[c])) for c in st_cols if int(r[c]) > 0), key=lambda kv: -kv[1])[:2]
        lines.append(f"  {100 * int(r[iS]) / tot:5.1f}%  exec={r[iEx]:>10}  {r[iSrc].strip()[:70]}  {main}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
