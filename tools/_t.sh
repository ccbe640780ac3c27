for n in 250 300 350 400 450 500; do echo "baseline $n: $(./tools/rt_sessions.bin $n 10 2>&1 | tail -2 | tr '\n' ' ')"; done
for n in 125 150 175 200; do echo "main $n: $(./tools/rt_sessions.bin $n 10 1920 1080 30 4000000 0 1 1 0 2>&1 | tail -2 | tr '\n' ' ')"; done
for n in 125 150 175; do echo "high $n: $(./tools/rt_sessions.bin $n 10 1920 1080 30 4000000 0 1 2 0 2>&1 | tail -2 | tr '\n' ' ')"; done
