python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -2 gpurun_out/r02_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("N=8 value", d["value"], "e2e", e["value"], "pinned", (e.get("pinned_input") or {}).get("value"), (e.get("pinned_input") or {}).get("error"), "h2d", {k:v for k,v in (e.get("h2d_ceiling") or {}).items() if k!='note'}, "rt", d.get("realtime",{}).get("late_frames"), d.get("realtime",{}).get("latency_ms"), d.get("realtime",{}).get("sessions"))
PY
