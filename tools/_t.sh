python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed|Error" | head
python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['kernel_ms'])"
