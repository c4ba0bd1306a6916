"""profiles/r0N_ncu_traffic.json from `ncu --set full` reports of one decode step (band kernels in layer order + the tail):
  python tools/ncu_traffic.py gpurun_out/r02_full_tc.ncu-rep gpurun_out/r02_full_tail.ncu-rep > profiles/r02_ncu_traffic.json"""
import csv, io, json, subprocess, sys
LABELS = ["hyper_synthesis.layer_0", "hyper_synthesis.layer_1", "hyper_synthesis.layer_2", "synthesis.base_conv+activation", "synthesis.out_conv"]


def rows(path):
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  r = list(csv.reader(io.StringIO(out)))
  hdr, units = r[0], r[1]
  for row in r[2:]:
    yield {h: (v, u) for h, v, u in zip(hdr, row, units)}


def mb(cell):
  v, u = cell
  f = float(v.replace(",", ""))
  return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


kernels = {}
i = 0
for path in sys.argv[1:]:
  for r in rows(path):
    rd, wr = mb(r["dram__bytes_read.sum"]), mb(r["dram__bytes_write.sum"])
    e = dict(dram_bytes_per_launch=int(rd + wr), read=int(rd), write=int(wr), us=float(r["gpu__time_duration.sum"][0]),
             kernel=r["Kernel Name"][0][:60])
    k = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    if k in r:
      e["tensor_pipe_active_pct_elapsed"] = float(r[k][0])
    kernels[LABELS[i]] = e
    i += 1
print(json.dumps(dict(source="ncu --set full --clock-control none under `python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-io-stage --no-side`: "
                             "the four band_gemm_tc_kernel<2> launches of the first timed step (-k regex:band_gemm_tc_kernel -s 12 -c 4) and one "
                             "tail launch (tail_tz_kernel since v3); summaries in profiles/r02_ncu_full_tc*.txt / r02_ncu_full_tail*.txt",
                      batch=24, step_total_dram_bytes=sum(v["dram_bytes_per_launch"] for v in kernels.values()), kernels=kernels), indent=1))
