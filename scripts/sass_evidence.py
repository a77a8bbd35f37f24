"""Write profiles/<tag>_sass_histogram.txt and profiles/<tag>_ptxas.txt: per kernel of libadmm_lstm_b200.so the opcode
histogram (tcgen05 -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, prefetch -> CCTL) and registers / spills /
shared memory from `nvcc -Xptxas -v` (runs on the build box: no GPU needed).

    python scripts/sass_evidence.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
lib = os.path.join(ROOT, "admm_lstm_b200", "libadmm_lstm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1).split(".")[0]] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "CCTL", "SYNCS", "MUFU", "HMMA", "FFMA", "LDG", "STG",
       "LDS", "ATOMS", "RED", "ATOM")
out = [f"# SASS opcode histogram per kernel of admm_lstm_b200/libadmm_lstm_b200.so (cuobjdump -sass, sm_100a)",
       "# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier, CCTL = prefetch.L2"]
for name, c in kernels.items():
    total = sum(c.values())
    dn = demangle(name)
    dn = re.sub(r"admm::\(anonymous namespace\)::", "", dn)
    out.append(f"\n{dn[:150]}\n  instructions {total}: " + ", ".join(f"{k} {c[k]}" for k in KEY if c[k]))
    out.append("  top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(12)))
open(os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt"), "w").write("\n".join(out) + "\n")

flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"]
lines = [f"# nvcc {' '.join(flags)} -c <file>: registers, spills, shared memory per kernel"]
csrc = os.path.join(ROOT, "admm_lstm_b200", "csrc")
for f in sorted(os.listdir(csrc)):
    if not f.endswith(".cu"):
        continue
    r = subprocess.run(["nvcc", *flags, "-c", os.path.join(csrc, f), "-o", "/tmp/_sass_ev.o"], capture_output=True, text=True)
    name = None
    for ln in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            name = re.sub(r"admm::\(anonymous namespace\)::", "", demangle(m.group(1)))[:130]
        m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m2:
            spill = m2.groups()
        m3 = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", ln)
        if m3 and name:
            lines.append(f"{f:22s} regs {m3.group(1):>3s}  spill st/ld {spill[1]}/{spill[2]}  static smem {m3.group(2) or 0:>6}  {name}")
            name = None
open(os.path.join(ROOT, "profiles", f"{tag}_ptxas.txt"), "w").write("\n".join(lines) + "\n")
print("wrote", f"profiles/{tag}_sass_histogram.txt", f"profiles/{tag}_ptxas.txt")
