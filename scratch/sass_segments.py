#!/usr/bin/env python
"""Static SASS statistics of the headline kernel between its barriers (no GPU needed): instructions, FP64, moves, local-memory
accesses per segment.  usage: sass_segments.py [object file]"""
import re, subprocess, sys, collections
obj = sys.argv[1] if len(sys.argv) > 1 else "fest-3d_b200/csrc/fused.o"
fun = "_ZN3f3d2g47k_fusedILi7ELi1ELi2ELb1ELb0EEEvNS_6ParamsENS_5KArgsENS0_5TMapsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append(m.group(2).strip())
segs, cur = [], []
for t in ins:
    cur.append(t)
    if "BAR.SYNC" in t:
        segs.append(cur); cur = []
segs.append(cur)
def op(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]
print("total", len(ins))
for n, sg in enumerate(segs):
    c = collections.Counter(op(t).split(".")[0] for t in sg)
    mov = sum(1 for t in sg if "IMAD.MOV" in t or re.match(r"(@!?P\d+\s+)?MOV ", t))
    print("seg %d (ends %s): n=%d fp64=%d mov=%d LDL=%d STL=%d LDS=%d STS=%d LDG=%d SEL=%d ISETP=%d BRA=%d CS2R=%d LDTM=%d STTM=%d" % (
        n, sg[-1][:40] if sg else "", len(sg), c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"], mov, c["LDL"], c["STL"], c["LDS"], c["STS"], c["LDG"], c["SEL"] + c["FSEL"], c["ISETP"], c["BRA"], c["CS2R"], c["LDTM"], c["STTM"]))
