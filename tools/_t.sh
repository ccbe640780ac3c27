python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("N=1 value", d["value"], "e2e", e["value"], "pinned", e.get("pinned_input"), "h2d", {k:v for k,v in (e.get("h2d_ceiling") or {}).items() if k!='note'}, "rt", d.get("realtime",{}).get("late_frames"), d.get("realtime",{}).get("latency_ms"), d.get("realtime",{}).get("sessions"))
PY
