#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the
built library: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA loads /
stores), HMMA (legacy mma.sync), FFMA.  Writes profiles/sass_summary.txt.   python scripts/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal_sequencing_b200", "libmsq_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "FFMA", "MUFU", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("msq::", "")
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_]*(?:\.[A-Za-z0-9_.]+)?)[\s;]", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in KEYS:
            cur[base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
    rows = [(n, c) for n, c in counts.items() if any(c[k] for k in KEYS[:7])]
    total = collections.Counter()
    lines = ["SASS mnemonic counts per kernel of libmsq_b200.so (cuobjdump -sass, sm_100a); kernels without tensor / TMA / "
             "TMEM instructions are summed in the last line", "%-78s " % "kernel" + " ".join("%12s" % k for k in KEYS)]
    for n, c in rows:
        lines.append("%-78s " % n[:78] + " ".join("%12d" % c[k] for k in KEYS))
    for n, c in counts.items():
        total.update(c)
    rest = collections.Counter()
    for n, c in counts.items():
        if not any(c[k] for k in KEYS[:7]):
            rest.update(c)
    lines.append("%-78s " % ("(%d other kernels: elementwise, pooling, decode select, SGEMM, optimizer ...)" % (len(counts) - len(rows))) +
                 " ".join("%12d" % rest[k] for k in KEYS))
    lines.append("%-78s " % "TOTAL" + " ".join("%12d" % total[k] for k in KEYS))
    txt = "\n".join(lines) + "\n"
    open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
