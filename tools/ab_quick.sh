run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log; python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('$1','step %.2f parse %.2f sort %.2f cover %.2f phase1 %.2f chain %.2f e2e %.1f M/s'%(j['ms_per_step'],s['ms_parse'],s['ms_sort'],s['ms_cover'],s['ms_phase1'],s['ms_chain'],j['e2e']['value']/1e6))"; }
run "$1"
