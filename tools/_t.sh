run() { python bench.py --no-cpu --no-e2e --workload $1 $3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', d['value'], d['ms_per_step'])"; }
run 1080p-main g1 "--groups 1"
run 1080p-main g2 "--groups 2"
run 1080p-main g4 "--groups 4"
run 1080p g2 "--groups 2"
run 1080p g4 "--groups 4"
