"""Instruction-mnemonic counts per kernel of the shipped library: the SASS evidence kept under profiles/.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hoomd-tf_b200", "lib", "libhtf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "LDSM", "MUFU.TANH",
         "MUFU.EX2", "MUFU.RSQ", "MUFU.RCP", "MUFU.SQRT", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HMUL2", "ATOMS", "ATOMG", "REDG",
         "RED.E", "LDGSTS", "LDS.128", "STG.E.128", "LDG.E.128", "SHFL", "VOTE", "BAR.SYNC", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS",
         "ELECT", "CCTL")
kern, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*$", "", kern)
        counts[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
print("SASS mnemonic counts per kernel of hoomd-tf_b200/lib/libhtf_b200.so (cuobjdump -sass, sm_100a); static counts")
print("tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, mbarrier tx -> SYNCS, mma.sync -> HMMA,")
print("ldmatrix -> LDSM, packed fp32 -> FFMA2/FMUL2/FADD2, system-scope flags -> *.STRONG.SYS\n")
for k in sorted(counts):
    if total[k] == 0:
        continue
    sel = ", ".join("%s x%d" % (w, c) for w, c in sorted(counts[k].items(), key=lambda x: -x[1]) if c)
    print("%-70s %6d instr | %s" % (k[:70], total[k], sel))
