# end-of-round capture: every command first runs to exit 0 without ncu; a number printed under ncu is never a bench value
set -x
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02n.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02n.log 2>&1
python bench.py > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err
python bench.py --no-cpu --steps 100 --warmup 10 > gpurun_out/bench_long_r02n.json 2> gpurun_out/bench_long_r02n.err
python bench.py --no-cpu --workload 1080p-main > gpurun_out/bench_main_r02n.json 2> gpurun_out/bench_main_r02n.err
python bench.py --no-cpu --workload 1080p-high > gpurun_out/bench_high_r02n.json 2> gpurun_out/bench_high_r02n.err
python bench.py --no-cpu --workload single > gpurun_out/bench_single_r02n.json 2> gpurun_out/bench_single_r02n.err
python bench.py --no-cpu --workload 4k > gpurun_out/bench_4k_r02n.json 2> gpurun_out/bench_4k_r02n.err
python bench.py --no-cpu --workload rgba720 > gpurun_out/bench_rgba_r02n.json 2> gpurun_out/bench_rgba_r02n.err
B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu"
$B > gpurun_out/plain_n.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02n.csv $B > gpurun_out/ncu_ln.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 44 -c 17 -f -o gpurun_out/prof_r02n $B > gpurun_out/ncu_fn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
tail -2 gpurun_out/gpu_tests_r02n.log; tail -1 gpurun_out/smoke_r02n.log
