# phase-1 slice width (start offsets per block) on the C2 bench workload; run on the GPU box
for sl in 16384 12288 10240; do
SSB_CHAIN_SLICE=$sl python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log
python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('slice',$sl,'step %.2f phase1 %.2f chain %.2f'%(j['ms_per_step'],s['ms_phase1'],s['ms_chain']))"
done
