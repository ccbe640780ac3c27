set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01m.log 2>&1
python bench.py > gpurun_out/bench_r01m.json 2> gpurun_out/bench_r01m.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01m.json 2> gpurun_out/bench_ref_r01m.err
python bench.py --no-cpu --workload 1080p-main > gpurun_out/bench_main_r01m.json 2> gpurun_out/bench_main_r01m.err
python bench.py --no-cpu --workload 1080p-high > gpurun_out/bench_high_r01m.json 2> gpurun_out/bench_high_r01m.err
B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu"
$B > gpurun_out/plain_m.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01m.csv $B > gpurun_out/ncu_lm.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 41 -c 16 -f -o gpurun_out/prof_r01m $B > gpurun_out/ncu_fm.log 2>&1
H="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --workload 1080p-high"
$H > gpurun_out/plain_mh.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_high_r01m.csv $H > gpurun_out/ncu_lmh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:inter_t8 -c 1 -f -o gpurun_out/prof_high_r01m $H > gpurun_out/ncu_fmh.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
tail -c 400 gpurun_out/bench_r01m.json
