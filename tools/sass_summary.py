"""Static SASS instruction counts per kernel of libspike_b200.so -> markdown (profiles/rNN_sass_summary.md).
usage: sass_summary.py [lib] > profiles/r02_sass_summary.md   (runs anywhere: cuobjdump only)"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "spike_petsc_b200/lib/libspike_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["DMMA", "UBLKCP", "UBLKPF", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "DFMA", "LDL", "STL"]
cnt = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); cnt[name] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        cnt[name]["total"] += 1
        for c in cols:
            if op.startswith(c):
                cnt[name][c] += 1
demangle = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
print("# SASS summary (cuobjdump -sass %s, sm_100a)\n" % lib)
print("Instruction counts per kernel (static).  `DMMA` = `DMMA.8x8x4` (FP64 tensor core; FP64 has no tcgen05 kind, so `mma.sync` DMMA "
      "is the Blackwell tensor path for this dtype), `UBLKCP` = `cp.async.bulk` (TMA engine, 1-D bulk copies), `UBLKPF` = "
      "`cp.async.bulk.prefetch.L2`, `SYNCS` = mbarrier operations, `LDL/STL` = local-memory (spill) traffic.\n")
print("| kernel | " + " | ".join(cols) + " | total |")
print("|---|" + "---|" * (len(cols) + 1))
for (n, c), d in sorted(zip(cnt.items(), demangle), key=lambda t: t[1]):
    short = re.sub(r"\(.*$", "", d).replace("void ", "")
    print("| `%s` | " % short + " | ".join(str(c[k]) for k in cols) + " | %d |" % c["total"])
