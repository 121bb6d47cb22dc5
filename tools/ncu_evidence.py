#!/usr/bin/env python
"""profiles/<round>_ncu_evidence.json from `ncu --set full` captures: per kernel class the DRAM bytes of one launch and
the utilisation of the units that bound it.
usage: ncu_evidence.py [--out profiles/r2_ncu_evidence.json] class=report.ncu-rep[:kernel substring] ..."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "r1_ncu_evidence.json")
args = sys.argv[1:]
if args and args[0] == "--out":
    out_path = os.path.join(ROOT, args[1]) if not os.path.isabs(args[1]) else args[1]
    args = args[2:]
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in args:
    cls, rep = arg.split("=", 1)
    sub = None
    if ":" in rep:
        rep, sub = rep.split(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        if sub and sub not in name:
            continue
        g = lambda k: float(r[h.index(k)].replace(",", ""))
        unit = lambda k: rows[1][h.index(k)]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        dr = g("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
        dw = g("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
        dur = g("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[unit("gpu__time_duration.sum")]
        cyc = g("sm__cycles_elapsed.max")
        sms = 148
        wav = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        out[cls] = {
            "kernel": name.strip(), "report": os.path.relpath(rep, ROOT), "duration_ms_under_ncu": dur * 1e3,
            "dram_bytes_per_launch": dr + dw,
            "limiter": {
                "issue_slots_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "shared_memory_wavefront_pct": 100.0 * wav / sms / cyc,
                "shared_bank_conflict_wavefronts_pct": 100.0 * g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / max(wav, 1.0),
                "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "dram_pct": g("dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + g("dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
                "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "registers": g("launch__registers_per_thread"),
            },
        }
        break
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
