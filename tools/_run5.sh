timeout 600 python -m pytest tests/test_gpu_offpolicy.py -q -x 2>&1 | tail -5
timeout 160 python tools/profile_td3.py time 2>&1 | tail -5
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.err; python - <<'PY'
import json
for line in open('gpurun_out/bench_1gpu.json'):
    if line.startswith('{"metric"'):
        d = json.loads(line)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
        print("fused", d["roofline"]["ms_per_launch"], "scan large", d["roofline_scan_large"]["achieved"], d["roofline_scan_large"]["frac"])
        print("trpo", d["other_configs"]["config3_trpo"])
        print("td3", {k: v for k, v in d["other_configs"]["config4_td3"].items() if k.startswith(("ms", "train"))})
        print("cpu", d["cpu_baseline"]["value"])
PY
