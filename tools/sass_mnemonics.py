"""Mnemonic histogram of the shipped cubin (cuobjdump -sass libsntc.so), per kernel family: the evidence that the product path is
tcgen05 / TMEM / TMA code.  python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt   (no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "shallow_ntc_b200", "libsntc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "LDSM", "ELECT", "ACQBULK", "UCGABAR", "FFMA", "MUFU", "F2FP")
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
  m = re.search(r"Function : (\S+)", line)
  if m:
    cur = m.group(1)
    per[cur] = collections.Counter()
    continue
  m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
  if m and cur:
    per[cur][m.group(1)] += 1
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# cuobjdump -sass shallow_ntc_b200/libsntc.so   archs: {archs}   kernels: {len(per)}")
tot = collections.Counter()
for k, c in per.items():
  for op, n in c.items():
    tot[op] += n
print("\n## whole library: watched mnemonics (all variants summed by prefix)")
for w in WATCH:
  n = sum(v for op, v in tot.items() if op.startswith(w))
  variants = sorted({op for op in tot if op.startswith(w)})
  print(f"{w:12s} {n:7d}   {' '.join(variants[:12])}")
print("\n## per kernel (demangled name prefix): instructions, and the watched ones that occur")
for k, c in per.items():
  name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0][:90]
  hits = {w: sum(v for op, v in c.items() if op.startswith(w)) for w in WATCH}
  print(f"{name:92s} {sum(c.values()):7d}  " + " ".join(f"{w}={n}" for w, n in hits.items() if n))
