for v in 0 1 2 3 4 5; do
  SSB_EMIT_VARIANT=$v timeout 300 python bench.py --scale 0.5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.log
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));print('variant $v ms_emit',d['stages_ms_per_step']['ms_emit'],'frac',d['roofline']['frac'])"
done
