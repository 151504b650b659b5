"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the in-tree library (cuobjdump -sass).
Run on the build box (no GPU needed): python scripts/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quintnet_b200", "lib", "libquintnet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|UTMACCTL|LDTM|STTM|MEMBAR\.[A-Z.]+|SYNCS|REDG|RED|MULTIMEM|MUFU|UCGABAR_ARV|UCGABAR_WAIT|ATOMG|LDGMC|STGMC)[.\w]*")
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur is None: continue
    m = pat.search(line)
    if m:
        key = m.group(1)
        full = m.group(0)
        if key == "UTCHMMA" and ".2CTA" in full: key = "UTCHMMA.2CTA"
        if key == "UTCBAR" and ("2CTA" in full or "MULTICAST" in full): key = "UTCBAR.2CTA.MULTICAST"
        if key == "UTMALDG" and "2CTA" in full: key = "UTMALDG.2CTA"
        counts[cur][key] += 1
print("# SASS evidence per kernel (cuobjdump -sass quintnet_b200/lib/libquintnet_b200.so): Blackwell mnemonics")
print("# UTCHMMA = tcgen05.mma bf16, UTCQMMA = tcgen05.mma fp8, .2CTA = cta_group::2, UTMALDG/UTMASTG = TMA tensor load/store,")
print("# UBLKCP = 1-D bulk async copy, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, MEMBAR.SYS = system fences,")
print("# UCGABAR = cluster barrier, MULTIMEM = NVLS multicast ld_reduce/st")
for k in sorted(counts):
    if counts[k]:
        print(k)
        print("    " + ", ".join(f"{a}={b}" for a, b in sorted(counts[k].items())))
