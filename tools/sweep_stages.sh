# phase-1 stage count x group length (chunks per group) on the C2 bench workload; run on the GPU box
for st in 1 3 4 5 6; do for g in 8 16; do
SSB_P1_STAGES=$st SSB_CHAIN_GROUP=$g python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log
python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('stages',$st,'group',$g,'step %.2f phase1 %.2f chain %.2f'%(j['ms_per_step'],s['ms_phase1'],s['ms_chain']))"
done; done
