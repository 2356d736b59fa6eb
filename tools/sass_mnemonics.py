"""Per-kernel counts of the tensor-core / TMA / TMEM SASS instructions in the built library (runs without a GPU):
    python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt
UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UCGABAR = cluster barrier (CTA-pair kernels), USETMAXREG = setmaxnreg, SYNCS = mbarrier ops."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "controllora_b200" / "libcontrollora_b200.so"
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCATOMSWS|UCGABAR_ARV|UCGABAR_WAIT|USETMAXREG|SYNCS|ACQBULK|UTMACCTL|UTMACMDFLUSH)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, cur, i = collections.OrderedDict(), None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = names[i]
            i += 1
            short = re.sub(r"^void clb::", "", full)
            short = re.sub(r"\(.*$", "", short)
            cur = counts.setdefault(short, collections.Counter())
            continue
        if cur is None:
            continue
        for tok in PAT.findall(line):
            cur[tok] += 1
    print(f"cuobjdump -sass {LIB.relative_to(LIB.parent.parent)} : tensor-core / TMA / TMEM instruction counts per kernel (sm_100a)")
    print("UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st, "
          "UTCBAR = tcgen05.commit, UCGABAR = cluster barrier, USETMAXREG = setmaxnreg\n")
    tot = collections.Counter()
    for k in sorted(counts):
        c = counts[k]
        tot.update(c)
        if any(t.startswith(("UTC", "UTMA", "LDTM", "STTM")) for t in c):
            print(f"{k:<78} " + "  ".join(f"{t}:{n}" for t, n in sorted(c.items())))
    print("\nwhole library: " + "  ".join(f"{t}:{n}" for t, n in sorted(tot.items())))
    print(f"kernels in the library: {len(counts)}; with tensor-core / TMA / TMEM instructions: {sum(1 for c in counts.values() if any(t.startswith(('UTC', 'UTMA', 'LDTM', 'STTM')) for t in c))}")


if __name__ == "__main__":
    sys.exit(main())
