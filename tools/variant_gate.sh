#!/bin/bash
# tools/variant_gate.sh <tag> <variant libb200enc.so> <variant libb200enc_checked.so>
# One-call gate for a kernel variant built with other -D flags: (1) the whole GPU parity suite with the variant IN PLACE of the in-tree library (so the plugin,
# the shim and the tools load it too), (2) device-resident A/B bench lines variant / main / variant / main, (3) only if the suite is green: the ncu capture of
# one P step and the launch list with the variant (each after the same command ran to exit 0 without ncu), (4) the default bench line. Everything lands in
# gpurun_out/; the in-tree libraries are restored at the end. A number printed under ncu is never a bench value.
T=${1:-v2}; V=${2:-variants/libb200enc_v2.so}; VC=${3:-variants/libb200enc_v2_checked.so}
C=media_b200/csrc
mkdir -p gpurun_out variants/main
cp $C/libb200enc.so $C/libb200enc_checked.so variants/main/
use() { cp "$1" $C/libb200enc.so; cp "$2" $C/libb200enc_checked.so; }
stamp() { echo "[$(date +%H:%M:%S)] $*" | tee -a gpurun_out/gate_$T.txt; }
DEV="python bench.py --no-cpu --no-e2e"
stamp start; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/gate_$T.txt
use $V $VC
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$T.log 2>&1; RC=$?
stamp "variant pytest rc=$RC: $(tail -1 gpurun_out/gpu_tests_$T.log)"
for i in 1 2; do
    use $V $VC
    timeout 150 $DEV > gpurun_out/bench_dev_v${i}_$T.json 2> gpurun_out/bench_dev_v${i}_$T.err
    stamp "variant dev $i: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_dev_v${i}_$T.json').read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'].get('k_me_fine'), d['kernel_ms'].get('k_me_coarse'))" 2>&1 | tail -1)"
    use variants/main/libb200enc.so variants/main/libb200enc_checked.so
    timeout 150 $DEV > gpurun_out/bench_dev_m${i}_$T.json 2> gpurun_out/bench_dev_m${i}_$T.err
    stamp "main dev $i: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_dev_m${i}_$T.json').read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms'].get('k_me_fine'), d['kernel_ms'].get('k_me_coarse'))" 2>&1 | tail -1)"
done
if [ $RC -eq 0 ]; then
    use $V $VC
    B="python bench.py --steps 3 --warmup 3 --sessions 32 --groups 1 --no-cpu --no-e2e"
    timeout 120 $B > gpurun_out/plain_$T.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 46 -c 18 -f -o gpurun_out/prof_$T $B > gpurun_out/ncu_f_$T.log 2>&1
    stamp "ncu full rc=$?"
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $B > gpurun_out/ncu_l_$T.log 2>&1
    stamp "ncu launch list rc=$?"
    timeout 400 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
    stamp "default bench (variant): $(cut -c1-300 gpurun_out/bench_$T.json | tail -1)"
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
    stamp "smoke (variant): $(tail -1 gpurun_out/smoke_$T.log)"
else
    timeout 400 python bench.py > gpurun_out/bench_main_$T.json 2> gpurun_out/bench_main_$T.err
    stamp "default bench (main): $(cut -c1-300 gpurun_out/bench_main_$T.json | tail -1)"
fi
use variants/main/libb200enc.so variants/main/libb200enc_checked.so
stamp end
