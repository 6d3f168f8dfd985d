import re, subprocess, sys, collections
def count(obj, kern):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    # split by function
    out = {}
    cur = None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); out[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(2)
            out[cur][op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD") and len(op.split("."))>1 else "")] += 1
            out[cur]["TOTAL"] += 1
    for k, c in out.items():
        if kern in k:
            wide = sum(v for kk, v in c.items() if kk.startswith("IMAD.WIDE"))
            print("%-28s total=%6d IMAD.WIDE=%5d IMAD(other)=%5d IADD3=%5d SHF=%4d LOP3=%4d ISETP=%4d CALL=%3d" % (
                k[:28], c["TOTAL"], wide, sum(v for kk, v in c.items() if kk.startswith("IMAD") and not kk.startswith("IMAD.WIDE")),
                c["IADD3"], c["SHF"], c["LOP3"], c["ISETP"], c["CALL"]))
for tag in sys.argv[1:] or ("build",):
    print("==", tag)
    count("%s/inst_bn254_g1.o" % tag, "k_accumulate")
    count("%s/inst_bls12381_g1.o" % tag, "k_accumulate")
