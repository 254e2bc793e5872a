"""Opcode histogram per kernel of libmnb200.so (cuobjdump -sass): evidence of which hardware paths each kernel uses --
UTCHMMA / LDTM (tcgen05 + TMEM), UTMALDG / UTMASTG (TMA), HMMA (mma.sync), LDSM (ldmatrix), FFMA2 (packed fp32), SYNCS
(mbarrier).  Writes profiles/r2_sass_opcodes.json.   python scripts/sass_opcodes.py"""
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mnasnet-pytorch_b200", "mnb200", "libmnb200.so")
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "FFMA2",
       "FADD2", "FFMA", "SYNCS", "BAR", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "F2FP", "SHFL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for (mangled, cnt), name in zip(kernels.items(), names):
        name = re.sub(r"\((int|bool|unsigned int)\)", "", name)          # template arguments print as (int)3
        short = re.sub(r"\(.*", "", name).replace("void mnb::", "").replace("mnb::", "").replace("void ", "")
        d = {k: cnt[k] for k in KEY if cnt.get(k)}
        d["total"] = sum(cnt.values())
        res[short if short not in res else short + " #" + mangled[-6:]] = d
    tot = collections.Counter()
    for d in res.values():
        for k, v in d.items():
            tot[k] += v
    res = {"library_totals": {k: tot[k] for k in KEY + ("total",) if tot.get(k)}, "kernels": res}
    path = os.path.join(ROOT, "profiles", "r2_sass_opcodes.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res["library_totals"]))
    for k, d in res["kernels"].items():
        if any(x in d for x in ("UTMALDG", "UTCHMMA")) or "dw_mma" in k:
            print(k[:70], {x: d[x] for x in ("UTMALDG", "UTMASTG", "UTCHMMA", "LDTM", "HMMA", "LDSM", "FFMA2") if x in d})


if __name__ == "__main__":
    main()
