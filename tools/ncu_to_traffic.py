"""profiles/ncu_traffic.json from ncu captures: `--set full` raw CSV of the single-kernel launches (tools/prof_kernels.py) plus the
launch list of one step (share of the step).  bench.py reads the file for roofline.traffic / tensor-pipe figures and prints them
only next to a launch of the same shape (`l`).
usage: ncu_to_traffic.py RAW.csv LAUNCHES.csv WORKLOAD L > profiles/ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

raw, launches, workload, L = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, to=None):
    x = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0}.get(u, 1.0)
    return x * scale


# which captured kernel answers which bench key (first match wins; order of tools/prof_kernels.py launches)
keys = [("lstm_fwd", r"lstm_fwd_seq"), ("lstm_bwd", r"lstm_bwd_seq"), ("adam", r"adam_kernel"), ("softmax_ce", r"softmax_ce"),
        ("gather", r"gather_embed"), ("vocab_gemm", r"gemm2_bf16x3_kernel<1, 1>|gemm2_bf16x3_kernel<\(bool\)1, \(bool\)1>"), ("topk", r"beam_row_topk")]
out = {}
for r in data:
    name = r[col["Kernel Name"]]
    for k, pat in keys:
        if k in out or not re.search(pat, name):
            continue
        out[k] = {"kernel": re.sub(r"\(.*", "", name).replace("lrcn::", ""), "l": L,
                  "traffic_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                  "duration_us": val(r, "gpu__time_duration.sum"),
                  "tensor_pipe_pct_of_elapsed": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                  "tensor_pipe_pct_of_active": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
        break
# share of the step from the launch list
tot = collections.defaultdict(float)
for row in csv.DictReader([l for l in open(launches) if not l.startswith("==")]):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v
    tot[re.sub(r"\(.*", "", row["Kernel Name"]).replace("lrcn::", "").replace("void ", "")] += v
T = sum(tot.values())
for k, e in out.items():
    e["share_of_step"] = round(tot.get(e["kernel"].replace("void ", ""), 0.0) / T, 4) if T else None
print(json.dumps({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum and tensor-pipe activity per launch from `ncu --set full --clock-control none` "
                              "(tools/gpu_round2.sh, tools/ncu_to_traffic.py); share_of_step from the ncu launch list of one step; read by bench.py",
                  workload: out}, indent=1))
