"""SASS mnemonic census of the shipped library: how often every Blackwell-specific instruction appears in each kernel.
Usage: python profiles/sass_census.py [path/to/libxfmr_b200.so] > profiles/r02_sass_census.txt
(`cuobjdump -sass` of the sm_100a cubin; UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA would be the legacy mma.sync path.)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "matrix-factorization-torch_b200/libxfmr_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
want = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "FADD2", "REDUX")
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in want:
            if op.startswith(w):
                kernels[cur][w] += 1
tot = collections.Counter()
print(f"{lib}: {len(kernels)} kernels")
print("counts per kernel: " + ", ".join(want))
for name, c in kernels.items():
    for w in want:
        tot[w] += c[w]
    if c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"]:
        short = demangle(name)
        short = re.sub(r"\(CUtensorMap_st.*", "", short)
        print(f"  {short[:88]:88s} instr {c['_total']:6d} | " + " ".join(f"{w}={c[w]}" for w in want if c[w]))
print("whole library: " + " ".join(f"{w}={tot[w]}" for w in want))
print("tensor-core kernels: %d of %d; legacy HMMA instructions: %d" % (sum(1 for c in kernels.values() if c["UTCHMMA"]), len(kernels), tot["HMMA"]))
