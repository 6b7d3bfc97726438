"""cuobjdump -sass libctl_b200.so | python tools/sass_evidence.py : per-kernel counts of the Blackwell tensor-core / TMA /
mbarrier SASS mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UBLKCP = cp.async.bulk[.tensor])."""
import collections, re, sys
cur, stats = None, collections.OrderedDict()
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA\.MULTICAST)?|UTCATOMSWS|UTMALDG\.\dD(?:\.2CTA)?|UTMASTG\.\dD|UTMACCTL|UBLKCP|LDTM|SYNCS)")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        stats[cur][m.group(1)] += 1
print("# per kernel: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store,")
print("# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; kernels not listed are plain SIMT kernels")
for k, c in stats.items():
    if any(n.startswith(("UTC", "UTMA", "UBLK", "LDTM")) for n in c):
        print(f"{k[:84]:84s} " + "  ".join(f"{n}={v}" for n, v in sorted(c.items())))
