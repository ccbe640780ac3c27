#!/bin/bash
# tools/bench_refresh.sh <tag> -- bench lines of the committed code only (no tests, no ncu): the default line first, then the other workloads
T=${1:-v2b}
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "default rc=$? $(cut -c1-120 gpurun_out/bench_$T.json | tail -1)"
for w in 1080p-main 1080p-high single 4k rgba720 portrait720; do
    timeout 120 python bench.py --no-cpu --workload $w > gpurun_out/bench_${w}_$T.json 2> gpurun_out/bench_${w}_$T.err; echo "$w rc=$? $(cut -c1-120 gpurun_out/bench_${w}_$T.json | tail -1)"
done
timeout 120 python bench.py --no-cpu --steps 100 --warmup 10 > gpurun_out/bench_long_$T.json 2> gpurun_out/bench_long_$T.err; echo "long rc=$?"
