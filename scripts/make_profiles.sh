#!/bin/bash
# Turn what scripts/gpu_bench_profile.sh <tag> brought back in gpurun_out/ into the tracked summaries under profiles/.
TAG=${1:-r01}
G=gpurun_out; P=profiles
cp $G/${TAG}_bench.json $G/${TAG}_bench_ref.json $G/${TAG}_launches.csv $G/${TAG}_gpu.csv $G/${TAG}_configs.jsonl $P/
python scripts/ncu_lines.py $G/${TAG}_fused.ncu-rep 60 > $P/${TAG}_fused_ncu_summary.txt
ncu -i $G/${TAG}_fused.ncu-rep --page details > $P/${TAG}_fused_ncu_details.txt
python scripts/ncu_lines.py $G/${TAG}_bgzf.ncu-rep 40 > $P/${TAG}_bgzf_ncu_summary.txt
ncu -i $G/${TAG}_bgzf.ncu-rep --page details > $P/${TAG}_bgzf_ncu_details.txt
python - <<PY
import csv, io, json, subprocess
def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
    def get(name):
        i = h.index(name); x = float(v[i]); unit = u[i]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return {"dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
            "duration_ms_under_ncu": float(v[h.index("gpu__time_duration.sum")])}
f = traffic("$G/${TAG}_fused.ncu-rep"); f["kernel"] = "k_fused"; f["dram_bytes_per_launch"] = f["dram_bytes_read"] + f["dram_bytes_write"]
f["fragments_per_launch"] = 10000000
f["source"] = "ncu --set full --clock-control none, profiles/${TAG}_fused_ncu_summary.txt"
f["command"] = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
json.dump(f, open("$P/fused_traffic.json", "w"), indent=1)
b = traffic("$G/${TAG}_bgzf.ncu-rep"); b["kernel"] = "k_bgzf_deflate"; b["dram_bytes_per_launch"] = b["dram_bytes_read"] + b["dram_bytes_write"]; b["input_bytes_per_launch"] = 1819477288
b["command"] = "python scripts/bgzf_probe.py"
json.dump(b, open("$P/bgzf_traffic.json", "w"), indent=1)
print(f); print(b)
PY
