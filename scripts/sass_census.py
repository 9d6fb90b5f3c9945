"""Mnemonic census of the fused likelihood kernels in the built library (cuobjdump -sass):
   python scripts/sass_census.py > profiles/r02_sass_census.txt
Shows that the hot path is tcgen05 / TMEM / bulk-copy code (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM =
tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier) and what the epilogue spends its instructions on."""
import collections, os, re, subprocess, sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scdef_b200", "libscdef_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, counts, regs = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"scdef::\(anonymous namespace\)::|scdef::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern][op.split(".")[0]] += 1
        counts[kern]["*"] += 1
keys = ["*", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU", "FFMA", "FMUL", "FADD", "F2FP", "PRMT", "STS", "LDS", "LDG", "STG", "STL", "LDL", "NANOSLEEP", "BAR"]
print("# cuobjdump -sass scdef_b200/libscdef_b200.so : static instruction counts per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier ops")
print(f"{'kernel':58s}" + "".join(f"{k:>9s}" for k in keys))
for k, c in counts.items():
    if "k_lik_flat" in k or "k_lik_tc" in k or "k_tc_" in k:
        print(f"{k[:58]:58s}" + "".join(f"{c.get(x, 0):9d}" for x in keys))
