"""Summarise an .ncu-rep (raw + source pages) into text: key counters per kernel and the top stall sites.
Usage: python profiles/ncu_summarize.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
traffic = []
for r in rows[2:]:
    # (the conditional fallback launches of the merged sweep exit at once: only launches that did work count)
    if ("sweep_kernel" in r[idx["Kernel Name"]] and "dram__bytes_read.sum" in idx
            and float(r[idx["gpu__time_duration.sum"]].replace(",", "")) > 20.0):
        def _bytes(name):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[name]]]
            return float(r[idx[name]].replace(",", "")) * scale
        traffic.append(_bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum"))
    print("====", r[idx["Kernel Name"]][:90])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    vals = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stall), reverse=True)[:8]
    print("  stalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in vals))
if traffic and len(sys.argv) > 3:   # third argument: where to write the per-launch DRAM traffic for bench.py
    import json
    json.dump({"sweep_dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic), "report": rep},
              open(sys.argv[3], "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for k in kern:
    h = k["hdr"]
    ix = {c: i for i, c in enumerate(h)}
    tot = sum(int(r[ix["# Samples"]]) for r in k["rows"]) or 1
    print("==== top stall sites:", k["name"][:80], "samples", tot)
    for r in sorted(k["rows"], key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        extra = {c.replace(" (Not Issued)", "!"): r[ix[c]] for c in h if c.startswith("stall_") and "Not Issued" in c and r[ix[c]] not in ("0", "")}
        print(f"  {100 * int(r[ix['# Samples']]) / tot:5.1f}% ex={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:58]:58s} {extra}")
