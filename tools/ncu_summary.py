"""Text summary of one kernel launch in an .ncu-rep (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [launch index] > profiles/rNN_<kernel>_ncu_full_summary.txt
Prints the metrics the roofline arithmetic needs (duration, DRAM bytes, tensor-pipe / issue utilisation, registers,
shared memory, occupancy) and the stall-reason totals of the source page (needs -lineinfo builds)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
row = data[which]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: launch {which} of {len(data)}")
print(f"kernel: {row[col['Kernel Name']]}   grid {row[col['Grid Size']]} block {row[col['Block Size']]}")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__inst_executed.sum"]
for w in WANT:
    for h in hdr:
        if h == w or h.endswith("." + w):
            print(f"{w:82s} {row[col[h]]:>18s} {units[col[h]]}")
            break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page holds one table per launch, separated by a "Kernel Name" line
tables, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = []
        tables.append(cur)
    elif cur is not None:
        cur.append(r)
if tables and which < len(tables) and len(tables[which]) > 2:
    t = tables[which]
    h = t[0]
    ix = {n: i for i, n in enumerate(h)}
    body = t[1:]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: sum(int(r[ix[n]] or 0) for r in body) for n in stalls}
    samples = sum(int(r[ix["# Samples"]] or 0) for r in body)
    execd = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
    print(f"\nwarp-state samples: {samples}; warp instructions executed: {execd}")
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {n:28s} {v:8d}  {100.0 * v / max(1, samples):5.1f} %")
    print("\nhottest instructions (samples, SASS, top stall):")
    for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        top = max(stalls, key=lambda n: int(r[ix[n]] or 0))
        print(f"  {int(r[ix['# Samples']] or 0):7d}  {r[ix['Source']].strip()[:72]:72s} {top}")
