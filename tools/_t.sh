python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed|Error" | head
python tools/frame_kernel_times.py A 2>&1 | head -2 | cut -c1-200
for n in 300 350; do echo "baseline $n: $(./tools/rt_sessions.bin $n 10 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-60,150-470)"; done
