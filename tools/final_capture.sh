# end-of-round capture: every command first runs to exit 0 without ncu; a number printed under ncu is never a bench value
T=${1:-r02n}
set -x
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$T.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --no-cpu --steps 100 --warmup 10 > gpurun_out/bench_long_$T.json 2> gpurun_out/bench_long_$T.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$T.json 2> gpurun_out/bench_reference_$T.err
for w in 1080p-main 1080p-high single 4k rgba720 portrait720; do python bench.py --no-cpu --workload $w > gpurun_out/bench_${w}_$T.json 2> gpurun_out/bench_${w}_$T.err; done
python tools/frame_kernel_times.py A > gpurun_out/frame_kernel_times_$T.txt 2>&1
for n in 250 300 350; do echo "baseline $n: $(./tools/rt_sessions.bin $n 10 2>&1 | tail -2 | tr '\n' ' ')"; done > gpurun_out/rt_sessions_$T.txt 2>&1
for n in 125 150 175; do echo "main $n: $(./tools/rt_sessions.bin $n 10 1920 1080 30 4000000 0 1 1 0 2>&1 | tail -2 | tr '\n' ' ')"; done >> gpurun_out/rt_sessions_$T.txt 2>&1
for n in 125 150; do echo "high $n: $(./tools/rt_sessions.bin $n 10 1920 1080 30 4000000 0 1 2 0 2>&1 | tail -2 | tr '\n' ' ')"; done >> gpurun_out/rt_sessions_$T.txt 2>&1
B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --no-e2e"
$B > gpurun_out/plain_n.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $B > gpurun_out/ncu_ln.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 46 -c 18 -f -o gpurun_out/prof_$T $B > gpurun_out/ncu_fn.log 2>&1
BM="$B --workload 1080p-main"
$BM > gpurun_out/plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_cabac --launch-skip 21 -c 7 -f -o gpurun_out/prof_cabac_$T $BM > gpurun_out/ncu_fm.log 2>&1
BR="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --no-e2e --workload rgba720"
$BR > gpurun_out/plain_r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ingest_rgba --launch-skip 3 -c 1 -f -o gpurun_out/prof_rgba_$T $BR > gpurun_out/ncu_fr.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
tail -2 gpurun_out/gpu_tests_$T.log; tail -1 gpurun_out/smoke_$T.log
