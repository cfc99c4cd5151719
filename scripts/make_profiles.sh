#!/bin/bash
# Turn what scripts/gpu_bench_profile.sh <tag> brought back in gpurun_out/ into the tracked summaries under profiles/.
TAG=${1:-r02}
G=gpurun_out; P=profiles
cp $G/${TAG}_bench.json $G/${TAG}_bench_ref.json $G/${TAG}_launches.csv $G/${TAG}_gpu.csv $P/
[ -f $G/${TAG}_c1_16m.json ] && cp $G/${TAG}_c1_16m.json $G/${TAG}_c1_launches.csv $P/
[ -f $G/${TAG}_pytest.log ] && tail -5 $G/${TAG}_pytest.log > $P/${TAG}_pytest_tail.txt
[ -f $G/${TAG}_smoke.log ] && tail -1 $G/${TAG}_smoke.log >> $P/${TAG}_pytest_tail.txt
[ -f $G/${TAG}_bgzf_probe.log ] && cat $G/${TAG}_bgzf_probe.log > $P/${TAG}_bgzf_probe.txt && [ -f $G/${TAG}_bgzf_probe_huff.log ] && (echo '-- GG_BGZF_MODE=huff (k_bgzf_deflate)'; cat $G/${TAG}_bgzf_probe_huff.log) >> $P/${TAG}_bgzf_probe.txt
[ -f $G/${TAG}_configs.jsonl ] && grep "^{" $G/${TAG}_configs.jsonl > $P/${TAG}_configs.jsonl
python scripts/ncu_lines.py $G/${TAG}_fused.ncu-rep 60 > $P/${TAG}_fused_ncu_summary.txt
python scripts/ncu_passes.py $G/${TAG}_fused.ncu-rep > $P/${TAG}_fused_ncu_passes.txt
ncu -i $G/${TAG}_fused.ncu-rep --page details > $P/${TAG}_fused_ncu_details.txt
if [ -f $G/${TAG}_deam.ncu-rep ]; then
  python scripts/ncu_lines.py $G/${TAG}_deam.ncu-rep 40 > $P/${TAG}_deam_ncu_summary.txt
  ncu -i $G/${TAG}_deam.ncu-rep --page details > $P/${TAG}_deam_ncu_details.txt
fi
python scripts/ncu_lines.py $G/${TAG}_bgzf.ncu-rep 40 > $P/${TAG}_bgzf_ncu_summary.txt
ncu -i $G/${TAG}_bgzf.ncu-rep --page details > $P/${TAG}_bgzf_ncu_details.txt
bash scripts/sass_mix.sh > $P/${TAG}_sass_mix.txt
python - <<PY
import csv, io, json, os, subprocess
def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
    def get(name):
        i = h.index(name); x = float(v[i]); unit = u[i]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return {"dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
            "duration_ms_under_ncu": float(v[h.index("gpu__time_duration.sum")]) / {"ns": 1e6, "us": 1e3, "ms": 1, "usecond": 1e3, "msecond": 1, "nsecond": 1e6}.get(u[h.index("gpu__time_duration.sum")], 1)}
f = traffic("$G/${TAG}_fused.ncu-rep"); f["kernel"] = "k_fused"; f["dram_bytes_per_launch"] = f["dram_bytes_read"] + f["dram_bytes_write"]
f["fragments_per_launch"] = 10000000
f["source"] = "ncu --set full --clock-control none, profiles/${TAG}_fused_ncu_summary.txt"
f["command"] = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
json.dump(f, open("$P/fused_traffic.json", "w"), indent=1)
b = traffic("$G/${TAG}_bgzf.ncu-rep"); b["kernel"] = "k_bgzf_lz"; b["dram_bytes_per_launch"] = b["dram_bytes_read"] + b["dram_bytes_write"]; b["input_bytes_per_launch"] = 1819477288
b["command"] = "python scripts/bgzf_probe.py"
json.dump(b, open("$P/bgzf_traffic.json", "w"), indent=1)
print(f); print(b)
if os.path.exists("$G/${TAG}_deam.ncu-rep"):
    d = traffic("$G/${TAG}_deam.ncu-rep"); d["kernel"] = "k_deam_stream"; d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]; d["records_per_launch"] = 16000000
    d["command"] = "python bench.py --config C1 --frags 16000000 --steps 2 --warmup 3 --no-cpu-baseline"
    json.dump(d, open("$P/deam_traffic.json", "w"), indent=1); print(d)
PY
