"""profiles/r02_sass_excerpts.txt: per kernel of libsige_b200.so, counts and first occurrences of the tensor-core / TMA /
TMEM / cluster SASS mnemonics (the proof that the hot kernels are Blackwell-native, B200_PROFILING.md).

    cuobjdump -sass sige_b200/lib/libsige_b200.so > /tmp/sass.txt && python tools/sass_excerpts.py /tmp/sass.txt > profiles/rNN_sass_excerpts.txt
"""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
funcs = re.split(r"\n\s*Function : ", txt)[1:]
print("# SASS evidence: cuobjdump -sass sige_b200/lib/libsige_b200.so (sm_100a), per kernel: instruction counts of the")
print("# tensor-core / TMA / TMEM / cluster mnemonics and the first occurrence of each.\n")
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCMXQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|HMMA|IMMA|LDSM|LDGSTS|UCGABAR|ACQBULK|ELECT|CCTL|REDG|STAS|MEMBAR|ERRBAR)(\.[A-Z0-9_.]+)?")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if "sige" not in name:
        continue
    cnt, first, n_ins = collections.Counter(), {}, 0
    for ln in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
        if not m:
            continue
        n_ins += 1
        mm = pat.search(m.group(1))
        if mm:
            cnt[mm.group(1)] += 1
            first.setdefault(mm.group(1), m.group(1).strip())
    if not cnt:
        continue
    print("## %s" % name[:200])
    print("   %d SASS instructions; %s" % (n_ins, ", ".join("%s x%d" % kv for kv in cnt.most_common())))
    for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDSM", "STAS", "UCGABAR"):
        if k in first:
            print("      %-8s e.g.  %s" % (k, first[k][:150]))
    print()
