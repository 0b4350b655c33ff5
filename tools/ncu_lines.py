"""Per-CUDA-source-line stall samples of one kernel in an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py x.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    elif hdr and r[0].isdigit() and len(r) == len(hdr):
        n = int(r[ix["# Samples"]] or 0)
        if n:
            st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:3]
            out.append((n, fname, int(r[0]), r[1].strip(), int(r[ix["Instructions Executed"]] or 0), st))
tot = sum(o[0] for o in out)
print(f"# {rep}: {tot} samples on {len(out)} source lines")
for n, f, ln, text, ex, st in sorted(out, reverse=True)[:top]:
    s = ", ".join(f"{b[6:]} {a}" for a, b in st if a)
    print(f"{n:6d} {100.0 * n / tot:5.1f}%  {f}:{ln:<4d} ex {ex:9d}  {text[:70]:70s} | {s}")
