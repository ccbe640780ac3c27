set -x
python bench.py > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01k.json 2> gpurun_out/bench_ref_r01k.err
python bench.py --no-cpu --workload 1080p-main > gpurun_out/bench_main_r01k.json 2> gpurun_out/bench_main_r01k.err
python bench.py --no-cpu --workload 1080p-main-1slice > gpurun_out/bench_main1_r01k.json 2> gpurun_out/bench_main1_r01k.err
B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu"
$B > gpurun_out/plain_k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01k.csv $B > gpurun_out/ncu_lk.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 41 -c 16 -f -o gpurun_out/prof_r01k $B > gpurun_out/ncu_fk.log 2>&1
M="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --workload 1080p-main"
$M > gpurun_out/plain_km.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_main_r01k.csv $M > gpurun_out/ncu_lkm.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 47 -c 18 -k regex:cabac -f -o gpurun_out/prof_main_r01k $M > gpurun_out/ncu_fkm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
tail -c 600 gpurun_out/bench_r01k.json
