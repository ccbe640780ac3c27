for w in 1080p-main 1080p-high single 4k rgba720 portrait720; do python bench.py --no-cpu --workload $w --steps 20 --realtime-seconds 0 2> gpurun_out/sv_$w.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$w', 'value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'dev', d['device_ms_per_step'])
print('   ', {k:v for k,v in d['kernel_ms'].items() if v>0.05})
"; done
python tools/frame_kernel_times.py 2>&1 | tail -12
