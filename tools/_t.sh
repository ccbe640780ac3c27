run() { python bench.py --no-cpu --workload 1080p-main --steps 20 --realtime-seconds 0 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$1 value', d['value'], 'e2e', e['value'], 'pinned', (e.get('pinned_input') or {}).get('value'))"; }
run base ""
B200ENC_CABAC_SLAB_MAXN=512 run slab100_all ""
B200ENC_CABAC_SLAB_MAXN=512 B200ENC_CABAC_SMEM_KB=60 run slab60_all ""
B200ENC_CABAC_SLAB_MAXN=512 run slab100_all_g8 "--groups 8"
run base_g8 "--groups 8"
