"""Prints the kernels of the last complete step in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/launch_list.py gpurun_out/launches.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
names = [r[ix["Kernel Name"]] for r in data]
vals = [float(r[ix["Metric Value"]].replace(",", "")) for r in data]
starts = [i for i, n in enumerate(names) if "prep_operand" in n and (i == 0 or "prep_operand" not in names[i - 1])]
s, e = starts[-2], starts[-1]
# bench.py flushes L2 between steps with an in-place add on a 256 MiB buffer: not part of the step
keep = [i for i in range(s, e) if "CUDAFunctorOnSelf_add" not in names[i]]
names, vals = [names[i] for i in keep], [vals[i] for i in keep]
s, e = 0, len(names)
tot = sum(vals[s:e])
print(f"one step (forward + backward): {e - s} launches, {tot / 1e3:.1f} us under ncu (cold caches, serialised)")
sweeps = sum(v for n, v in zip(names[s:e], vals[s:e]) if ("sweep_kernel" in n or "wg_kernel" in n or "rt_kernel" in n))
print(f"sweep kernels: {sweeps / 1e3:.1f} us = {100 * sweeps / tot:.1f} % of the step")
for n, v in zip(names[s:e], vals[s:e]):
    print(f"  {v / 1e3:9.2f} us  {n[:120]}")
