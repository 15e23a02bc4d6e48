"""Per-kernel SASS census of liblossyless_b200.so: which kernels carry tcgen05 (UTC*MMA), TMEM loads (LDTM),
TMA (UTMALDG / UTMASTG / UBLKCP), legacy tensor-core (HMMA) and special-function (MUFU) instructions.

    python tools/sass_census.py > profiles/r02_sass_census.txt

Runs on the build box (cuobjdump only needs the .so, no GPU).  See /opt/skills/guides/B200_PROFILING.md,
"What proves a Blackwell-native kernel".
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lossyless_b200", "_C", "liblossyless_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDSM", "LDGSTS",
             "MUFU", "SYNCS", "ACQBULK", "PREEXIT", "ATOMG", "RED", "LDG.E.128", "STG.E.128"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            cur["__variants"] += 1
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["__instructions"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                cur[mn] += 1
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts summed over template variants")
    cols = [m for m in MNEMONICS if any(k[m] for k in kernels.values())]
    print("kernel".ljust(44) + "var".rjust(4) + "inst".rjust(8) + "".join(c.rjust(10) for c in cols))
    for name, c in kernels.items():
        short = name.replace("ll::", "")
        print(short[:43].ljust(44) + str(c["__variants"]).rjust(4) + str(c["__instructions"]).rjust(8) + "".join((str(c[m]) if c[m] else ".").rjust(10) for m in cols))
    tc = [n for n, c in kernels.items() if c["UTCHMMA"]]
    print(f"\n# kernels with tcgen05.mma (UTCHMMA): {', '.join(n.replace('ll::', '') for n in tc)}")
    print(f"# kernels with legacy mma.sync (HMMA): {', '.join(n.replace('ll::', '') for n, c in kernels.items() if c['HMMA'])}")


if __name__ == "__main__":
    sys.exit(main())
