L=media_b200/host/libVideoCodec.so
run() { echo "$1: $(env $2 ./tools/e2e_plugin.bin $L $3 30 2>/dev/null | tail -1 | cut -c100-330)"; }
run "N=128" "X=1" 128
run "N=192" "X=1" 192
run "N=256" "X=1" 256
run "N=256 max48" "B200ENC_BATCH_MAX=48" 256
run "N=256 max24" "B200ENC_BATCH_MAX=24" 256
for cfg in "128 4" "192 6" "256 8"; do set -- $cfg; python bench.py --no-cpu --no-e2e --steps 30 --sessions $1 --groups $2 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value $1/$2', d['value'], d['ms_per_step'])"; done
