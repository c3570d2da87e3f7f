"""Rebuild profiles/traffic.json from an `ncu --set full --page raw --csv` export of one step of the
headline config (bench.py --batch 20000 --steps 1 --warmup 3 --no-large; the seven launches of the
timed step: the reduction kernel and the six occupancy phases of the eigenvalue kernel).
Usage: python scripts/ncu_traffic.py profiles/r2_c2_final_kernels_ncu_full_raw.csv [problems]"""
import csv
import json
import sys

path = sys.argv[1]
problems = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, key, scale_unit=None):
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]]
    if scale_unit == "bytes":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if scale_unit == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    return v


def entry(r):
    return {
        "kernel": r[ix["Kernel Name"]],
        "ms": round(val(r, "gpu__time_duration.sum", "ms"), 3),
        "dram_read_GB": round(val(r, "dram__bytes_read.sum", "bytes") / 1e9, 4),
        "dram_write_GB": round(val(r, "dram__bytes_write.sum", "bytes") / 1e9, 4),
        "pipe_fp64_active_pct": round(val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 2),
        "issue_active_pct": round(val(r, "sm__issue_active.avg.pct_of_peak_sustained_active")
                                  if "sm__issue_active.avg.pct_of_peak_sustained_active" in ix
                                  else val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "warps_active_pct": round(val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
        "warp_instructions": int(val(r, "smsp__inst_executed.sum")),
        "smem_bank_conflicts": int(val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")),
        "registers": int(val(r, "launch__registers_per_thread")),
    }


eig = [entry(r) for r in data if "rpqr_eig32" in r[ix["Kernel Name"]]]
red = [entry(r) for r in data if "rphess_pair32" in r[ix["Kernel Name"]]]
assert eig and red, "expected the reduction kernel and the eigenvalue phases in the capture"
eig_bytes = sum((e["dram_read_GB"] + e["dram_write_GB"]) * 1e9 for e in eig)
red_bytes = sum((e["dram_read_GB"] + e["dram_write_GB"]) * 1e9 for e in red)
eig_ms = sum(e["ms"] for e in eig)
out = {
    "comment": "DRAM traffic and issue counters of the FINAL headline kernels from one ncu --set full capture "
               f"({path}: bench.py --batch {problems} --steps 1 --warmup 3 --no-large, the launches of the timed step), "
               "per launch divided by the problems of the launch; written by scripts/ncu_traffic.py",
    "problems_per_launch_in_capture": problems,
    "rphess_pair32_dram_bytes_per_problem": red_bytes / problems / len(red),
    "rpqr_eig32_dram_bytes_per_problem": eig_bytes / problems,
    "rpqr_eig32_phases": eig,
    "rphess_pair32": red[0],
    "warp_instructions_per_problem_iteration": sum(e["warp_instructions"] for e in eig) / problems,
    "pipe_fp64_active_pct_time_weighted": sum(e["pipe_fp64_active_pct"] * e["ms"] for e in eig) / eig_ms,
    "issue_active_pct_time_weighted": sum(e["issue_active_pct"] * e["ms"] for e in eig) / eig_ms,
    "warps_active_pct_time_weighted": sum(e["warps_active_pct"] * e["ms"] for e in eig) / eig_ms,
    "note": "the six occupancy phases hand the packed prefix of every problem to the next phase through HBM, "
            "so the DRAM traffic of the iteration is a multiple of the 36.9 KB it needs once",
}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "rpqr_eig32_phases"}, indent=1)[:1500])
