"""SASS evidence per hot kernel (the .so is git-ignored, so this excerpt is the tracked record):
    python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt"""
import collections, datetime, re, subprocess
SO = "bodywork-mlops-demo_b200/libb2gram.so"
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {SO}  ({datetime.datetime.utcnow():%Y-%m-%dT%H:%MZ})")
print("# mnemonic counts over the whole library (B200_PROFILING.md: UTC*MMA = tcgen05.mma, UTMALDG / UBLKCP = TMA, LDTM / STTM = tcgen05.ld / st)")
for m in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "LDSM", "FHFMA.BF16", "HFMA2.BF16_V2", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "DMMA", "MUFU.RCP64H", "MEMBAR.SC.SYS", "NANOSLEEP"):
    print(f"{m:34s} {sass.count(m)}")
funcs = re.split(r"(?=\s+Function : )", sass)
WANT = r"LDSM|FHFMA|HFMA2\.BF16|HADD2\.BF16|ELECT|UTCHMMA|UTMALDG|UBLKCP|LDTM|STTM|UTCBAR|SYNCS\.|DMMA|MUFU\.RCP64H|FFMA2|\.SYS|MEMBAR|ATOMG|REDG|NANOSLEEP|UTCATOM|DFMA|LDS\.128|STS\.128|BAR\.SYNC|ST\.E|LDG"
for pat in ("gram_tc_kernelIfLi128ELb1", "gram_b16_split_kernel", "gram_b16_single_kernel", "tc_shift_kernelIf", "tc_finalize_kernel", "solve_cholesky_kernel", "solve_eigvals_kernel",
            "score_narrow_kernelIfLi1ELb1ELb1", "gram_narrow_kernelIfLi8", "p2p_scatter_kernel", "p2p_gather_kernel"):
    body = next((f for f in funcs if re.search(r"Function : \S*" + pat, f)), None)
    if body is None:
        continue
    name = re.search(r"Function : (\S+)", body).group(1)
    name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:160]
    lines = [l for l in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    print(f"\n## {name}\n# {len(lines)} instructions; per opcode of interest: count, first two occurrences")
    seen = collections.OrderedDict()
    for l in lines:
        txt = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip()
        m = re.match(r"/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?(\S+)", txt)
        if not m or not re.search(WANT, m.group(3)):
            continue
        seen.setdefault(m.group(3), []).append(txt)
    for op, occ in sorted(seen.items()):
        print(f"  {op:36s} x{len(occ)}")
        for t in occ[:2]:
            print(f"      {t}")
