"""Opcode histogram of the hot kernels from the built objects (cuobjdump -sass): the Blackwell-native evidence
(UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UBLKCP = bulk copy, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDGSTS = cp.async, VIMNMX / VIMNMX3 = the shortlist networks, FFMA2 / FMUL2 / FADD2 = packed
fp32, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents) kept under profiles/ so that it
survives a rebuild.   usage: python scripts/sass_report.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "superpoint-stereo-visual-odometry_b200", "csrc")
WANT = {"match_tc.o": ["k_tc_gemm", "k_tc_fallback", "k_tc_triage", "k_tc_rerank"],
        "decode.o": ["k_softmax_heat", "k_detect", "k_desc_planes", "k_desc_normalize"]}
for obj, kernels in WANT.items():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    fn, ops = None, {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            ops[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and fn:
            ops[fn][m.group(2)] += 1
    for fn, c in ops.items():
        dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        if not any(k in dem for k in kernels):
            continue
        short = re.split(r"\((?:CUtensorMap|const|int|float|spvo::DetectParams)", dem)[0].replace("void ", "").replace("spvo::", "")
        print(f"== {short}   [{obj}]  {sum(c.values())} SASS instructions")
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        print("   " + ", ".join(f"{op} {n}" for op, n in base.most_common(18)))
        special = {op: n for op, n in c.items() if re.match(r"(UTCHMMA|UTMALDG|LDTM|STTM|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS|LDGSTS|VIMNMX3|UCGABAR|MEMBAR|ATOMS|RED|REDUX|FFMA2|FMUL2|FADD2|ACQBULK|PREEXIT)", op)}
        if special:
            print("   evidence: " + ", ".join(f"{op} x{n}" for op, n in sorted(special.items())))
