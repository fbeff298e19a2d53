#!/usr/bin/env python
"""SASS opcode histogram of the shipped libtcfd.so (cuobjdump -sass), per kernel family: the evidence that the
binary uses TMA / bulk copies / mbarriers / packed f32x2 math / tcgen05 (profiles/*_sass_histogram.txt).
usage: python scripts/sass_histogram.py [path/to/libtcfd.so] > profiles/<tag>_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "torch-cfd_b200", "libtcfd.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fam = collections.defaultdict(collections.Counter)
nker = collections.Counter()
cur = None
KEY = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR",
       "UTCATOMSWS", "HMMA", "LDGSTS", "BAR", "LDS", "STS", "LDG", "STG", "FFMA", "DFMA", "MUFU", "RED", "ATOMG", "TRAP", "NANOSLEEP")
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = m.group(1)
        for f in ("ns2d_flow_kernel", "ns2d_small_kernel", "ns2d_rows3_kernel", "ns2d_rows2_kernel", "ns2d_cols2_kernel", "ns2d_rows_kernel",
                  "ns2d_cols_kernel", "ns2d_record_kernel", "sconv_planes_fwd2", "sconv_planes_inv2", "sconv_planes_fwd_kernel",
                  "sconv_planes_inv_kernel", "sconv_xaxis", "sconv_mix", "fno_layer_glue_tc_kernel", "fno_layer_glue_kernel",
                  "fno_linear_kernel", "fno_project_kernel", "fft2_xaxis", "fft2_c2r", "fft2_r2c", "resample_bilinear"):
            if f in name:
                cur = f
                break
        else:
            cur = "other"
        nker[cur] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        op = m.group(1)
        fam[cur]["_total"] += 1
        for k in KEY:
            if op == k or op.startswith(k) and k in ("UTCATOMSWS",):
                fam[cur][k] += 1
print(f"SASS opcode histogram of {os.path.basename(lib)} ({os.path.getsize(lib)} bytes), cuobjdump -sass; instantiations per family in ()")
print("%-28s %9s " % ("kernel family", "instr") + " ".join("%8s" % k[:8] for k in KEY))
tot = collections.Counter()
for f in sorted(fam):
    print("%-28s %9d " % (f"{f} ({nker[f]})", fam[f]["_total"]) + " ".join("%8d" % fam[f][k] for k in KEY))
    tot.update(fam[f])
print("%-28s %9d " % ("ALL", tot["_total"]) + " ".join("%8d" % tot[k] for k in KEY))
