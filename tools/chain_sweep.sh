for cfg in "2 16384" "2 12288" "4 16384" "3 16384" "1 16384"; do
  set -- $cfg
  SSB_CHAIN_GROUP=$1 SSB_CHAIN_SLICE=$2 SSB_CHAIN_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.log
  echo "group=$1 slice=$2: $(grep 'chunks=' gpurun_out/sw.log | tail -1)"
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));print('   ms_chain',d['stages_ms_per_step']['ms_chain'],'total',d['ms_per_step'])"
done
