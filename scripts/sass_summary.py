"""Per-kernel SASS evidence of the Blackwell-native paths in libqt_b200.so (no GPU needed):
counts of UTC*MMA (tcgen05.mma; .2CTA = cta_group::2; with a tmem scale operand = block_scale), UTCCP (tcgen05.cp), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF (TMA load / store /
prefetch), SYNCS (mbarrier), USETMAXREG (setmaxnreg), HMMA (legacy mma.sync -- must be 0), and registers per kernel.
    python scripts/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantized-training_b200", "quantized_training", "_lib", "libqt_b200.so")
PAT = {"UTCMMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UTMAPF": r"\bUTMAPF", "SYNCS": r"\bSYNCS", "USETMAXREG": r"\bUSETMAXREG", "HMMA": r"\bHMMA", "LDGSTS": r"\bLDGSTS",
       "REDUX": r"\bREDUX", "MUFU.EX2": r"MUFU\.EX2", "UTCMMA.2CTA": r"\bUTC[A-Z]*MMA\.2CTA",
       "UTCMMA.blockscale": r"\bUTC[A-Z]*MMA gdesc.*idesc\[UR\d+\], tmem", "UTCCP": r"\bUTCCP", "UTMALDG.2CTA": r"\bUTMALDG\.\dD\.2CTA",
       "UTCBAR.MULTICAST": r"\bUTCBAR\.2CTA\.MULTICAST"}
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
demangled = dict(zip(counts, subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()))


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)
    return n[:150]


print("# cuobjdump -sass libqt_b200.so (sm_100a): instruction counts per kernel; kernels with tensor-core / TMA / TMEM code first")
print("# " + " ".join(f"{k:>10s}" for k in PAT) + "  regs  kernel")
rows = sorted(counts.items(), key=lambda kv: (-kv[1]["UTCMMA"], -kv[1]["UTMALDG"], kv[0]))
groups = collections.OrderedDict()
for name, c in rows:
    d = short(demangled.get(name, name))
    base = re.sub(r"<.*", "", d)
    if c["UTCMMA"] or c["UTMALDG"] or c["LDTM"]:
        print("  " + " ".join(f"{c[k]:10d}" for k in PAT) + f"  {regs.get(name, 0):4d}  {d}")
    else:
        g = groups.setdefault(base, collections.Counter())
        g["n"] += 1
        for k in PAT:
            g[k] += c[k]
print("# other kernel families (summed over their template instantiations)")
for base, g in groups.items():
    print("  " + " ".join(f"{g[k]:10d}" for k in PAT) + f"  {g['n']:4d}x {base}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# totals: " + ", ".join(f"{k}={tot[k]}" for k in PAT) + f"; kernels: {len(counts)}")
assert tot["HMMA"] == 0, "legacy mma.sync code present"
