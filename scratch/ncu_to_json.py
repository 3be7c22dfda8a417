"""Fold an .ncu-rep into profiles/ncu_r2.json (the measured per-launch DRAM traffic / tensor-pipe numbers bench.py quotes):
   python scratch/ncu_to_json.py <rep> <units per launch (particles, or 0)> <label> [kernel-regex]"""
import csv, json, os, re, subprocess, sys
rep, units, label = sys.argv[1], float(sys.argv[2]), sys.argv[3]
pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
col = {h: i for i, h in enumerate(H)}
def val(r, k, scale=None):
    if k not in col or r[col[k]] in ("", "n/a"):
        return None
    v, u = float(r[col[k]]), U[col[k]]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * mult
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_r2.json")
out = json.load(open(path)) if os.path.exists(path) else {}
seen = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    key = re.sub(r"^void ", "", name).split("<")[0].split("(")[0]
    seen.setdefault(key, []).append(r)
for key, rs in seen.items():
    r = rs[-1]                                             # last captured launch of the kernel
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    e = {"kernel_name": r[col["Kernel Name"]], "source": f"profiles/{label} (ncu --set full --clock-control none, {os.path.basename(rep)})",
         "units_per_launch": units, "duration_s_under_ncu": val(r, "gpu__time_duration.sum"),
         "dram_bytes_per_launch": None if rd is None else rd + wr,
         "dram_bytes_per_particle": None if rd is None or not units else (rd + wr) / units,
         "registers_per_thread": val(r, "launch__registers_per_thread"),
         "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warp_instructions": val(r, "smsp__inst_executed.sum"),
         "thread_instructions_per_particle": None if not units else 32.0 * val(r, "smsp__inst_executed.sum") / units,
         "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
         "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
         "fma_pipe_cycles_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")}
    out[key] = e
    print(key, json.dumps(e)[:300])
json.dump(out, open(path, "w"), indent=1)
