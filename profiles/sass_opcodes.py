"""Opcode histogram of the shipped library: python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.txt

Static instruction counts per kernel from `cuobjdump -sass`, restricted to the opcodes that prove the Blackwell path
(tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, mbarrier -> SYNCS,
cluster barrier -> UCGABAR_*) plus atomics; HMMA (legacy mma.sync) must be absent."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ieee_b200", "libieee_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEEP = re.compile(r"^(UTC\w+|LDTM|STTM|UTMA\w+|SYNCS|UCGABAR_\w+|HMMA|IMMA|ATOMS|ATOMG|REDG|RED|UBLKCP|UTCATOMSWS)$")
print("# SASS opcode evidence for ieee_b200/libieee_b200.so (cuobjdump -sass; architectures in the file: %s)" % ", ".join(archs))
print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, TMA loads/stores -> UTMALDG / UTMASTG, mbarrier -> SYNCS,")
print("# cluster barrier -> UCGABAR_*; HMMA (legacy mma.sync) must be absent.  Counts are static instruction counts per kernel.")
print()
hmma = 0
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
    name, body = demangle(m.group(1)), m.group(2)
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", body, re.M)
    hist = collections.Counter(o for o in ops if KEEP.match(o))
    hmma += hist.get("HMMA", 0)
    short = re.sub(r"\(.*", "", name)
    print("%-70s %6d instr  %s" % (short[:70], len(ops), "  ".join("%s=%d" % kv for kv in sorted(hist.items()))))
print()
print("# HMMA instructions in the whole library: %d" % hmma)
