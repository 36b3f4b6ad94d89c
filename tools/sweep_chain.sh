for g in 4 8 16 32; do for sl in 8192 16384 32768; do
SSB_CHAIN_GROUP=$g SSB_CHAIN_SLICE=$sl python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log
python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('group',$g,'slice',$sl,'step %.2f phase1 %.2f chain %.2f'%(j['ms_per_step'],s['ms_phase1'],s['ms_chain']))"
done; done
