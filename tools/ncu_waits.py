"""Per-launch summary of where the warps of the persistent layer kernel wait (ncu source page, SASS level):
python tools/ncu_waits.py file.ncu-rep.  Prints, per launch, the sampled stall counts on every mbarrier
try-wait, the cluster barrier and the MMA / TMA issue instructions."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for k in range(len(secs) - 1):
  s, e = secs[k], secs[k + 1]
  hdr = rows[s + 1]
  ci = {h: i for i, h in enumerate(hdr)}
  body = [r for r in rows[s + 2:e] if len(r) > 10]
  S, E = ci["# Samples"], ci["Instructions Executed"]
  tot = sum(int(r[S]) for r in body)
  key = (rows[s][1][:60], tot, len(body))
  if key in seen:
    continue
  seen.add(key)
  print(f"== launch section {k}: {rows[s][1][:70]}  instrs {len(body)}  samples {tot}")
  for i, r in enumerate(body):
    t = r[1].strip()
    hot = int(r[S]) > 0.01 * tot
    if "TRYWAIT" in t or "UCGABAR" in t or "UTCHMMA" in t or "UTMALDG" in t or "UTCBAR" in t or hot:
      if int(r[E]) > 0 and (hot or "TRYWAIT" in t):
        print(f"   {i:6d} {t[:80]:80s} samples {int(r[S]):7d} exec {int(r[E]):9d}")
