"""SASS evidence for profiles/: per-kernel counts and sample lines of the tcgen05 / TMEM / TMA mnemonics in libfibinet_b200.so.

    python tools/sass_evidence.py > profiles/r2_sass_tcgen05.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ctr_recommendation_b200", "libfibinet_b200.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|LDTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|UCGABAR_\w+)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    per, sample, cur = collections.OrderedDict(), {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            per[cur] = collections.Counter()
            continue
        m = PAT.search(line)
        if m and cur and "/*" in line:
            per[cur][m.group(1)] += 1
            sample.setdefault((cur, m.group(1)), line.rstrip())
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("# SASS evidence that the GEMM path is tcgen05 / TMEM / TMA (cuobjdump -sass ctr_recommendation_b200/libfibinet_b200.so, sm_100a, CUDA 12.9)")
    print("# tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2); tcgen05.ld -> LDTM; cp.async.bulk.tensor -> UTMALDG; tcgen05.commit -> UTCBAR")
    print("# kernel template argument 1 = precision: 1 tf32x3 (kind::tf32), 2 bf16, 3 tf32x2, 4 f16x3 (kind::f16 on fp16 hi|lo operands)")
    print("# per kernel: mnemonic counts, then one sample instruction line per mnemonic\n")
    print("TOTAL over the library: " + ", ".join(f"{k} x{v}" for k, v in sorted(total.items())) + "\n")
    for name, c in per.items():
        if not any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for k in c):
            continue
        print(name)
        print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
        for k in ("LDTM.x32", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.2D.2CTA"):
            if (name, k) in sample:
                print("  " + sample[(name, k)])


if __name__ == "__main__":
    sys.exit(main())
