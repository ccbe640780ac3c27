python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed|Error" | head -20
B200ENC_TRACE=1 ./tools/rt_sessions.bin 20 3 1920 1080 30 4000000 0 1 1 0 2>&1 | grep "coded twice" | tail -8
for n in 125 150 200; do echo "main auto $n: $(./tools/rt_sessions.bin $n 8 1920 1080 30 4000000 0 1 1 0 2>&1 | tail -2 | cut -c1-60,140-420 | tr '\n' ' ')"; done
