# A/B of the tokeniser's register budget on the GPU box: 5 blocks/SM (96 registers) against 4 (112)
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log; python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('$1','step %.2f parse %.2f phase1 %.2f chain %.2f e2e %.1f M/s'%(j['ms_per_step'],s['ms_parse'],s['ms_phase1'],s['ms_chain'],j['e2e']['value']/1e6))"; }
run minblocks5
touch stochasticsim_b200/csrc/spike.cu; make EXTRA_NVFLAGS=-DSSB_PARSE_MINBLOCKS=4 > gpurun_out/remake.log 2>&1; run minblocks4
