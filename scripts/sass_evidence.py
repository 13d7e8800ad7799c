#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native instructions in libavatarcraft_b200.so (cuobjdump -sass; no GPU needed):
counts of UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA), RED/ATOM, HMMA (legacy mma.sync -- must be 0),
plus the first line of each kind.  Writes profiles/sass_<tag>.txt.    python scripts/sass_evidence.py r02"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "avatarcraft_b200", "libavatarcraft_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KINDS = [("UTCHMMA", r"\bUTCHMMA"), ("UTC*MMA other", r"\bUTC(?!HMMA|BAR)[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
         ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("REDG/RED", r"\bREDG?\."), ("ATOMG", r"\bATOMG"),
         ("MUFU", r"\bMUFU"), ("HMMA (legacy)", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS")]
cur, per, first = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        per.setdefault(cur, collections.Counter()); first.setdefault(cur, {})
        continue
    if cur is None:
        continue
    for name, pat in KINDS:
        if re.search(pat, line):
            per[cur][name] += 1
            first[cur].setdefault(name, re.sub(r"\s+", " ", line.split("*/")[1] if "*/" in line else line).strip()[:110])
with open(os.path.join(ROOT, "profiles", f"sass_{tag}.txt"), "w") as f:
    f.write(f"# SASS evidence, libavatarcraft_b200.so ({tag}): cuobjdump -sass, instruction counts per kernel (static), first occurrence quoted\n")
    f.write("# arch: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))) + "\n\n")
    for k, c in per.items():
        if not c:
            continue
        f.write(f"{k}\n    " + "  ".join(f"{n}={v}" for n, v in c.items()) + "\n")
        for n in ("UTCHMMA", "LDTM", "UTMALDG", "REDG/RED"):
            if n in first[k]:
                f.write(f"      {n}: {first[k][n]}\n")
        f.write("\n")
    legacy = sum(c["HMMA (legacy)"] for c in per.values())
    f.write(f"# kernels: {len(per)}; legacy HMMA instructions in the whole library: {legacy}\n")
print(open(os.path.join(ROOT, "profiles", f"sass_{tag}.txt")).read()[:3000])
