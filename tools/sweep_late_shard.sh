# phase 1 of a shard far down the stream, emulated on one GPU: SSB_CHAIN_EXTRA_VAR = variance of the draws of k C2 contigs in front (4/9 per locus);
# group = chunks per phase-1 group (auto: the library's choice).  Round 2, B200 (ms): k=1: auto 14.5, 8 17.8, 16 15.0, 24 15.8, 32 19.6;
# k=3: 8 24.2, 16 20.2, 24 18.8, 32 22.0, 48 27.3;  k=7: 8 31.1, 16 27.8, 24 24.4, 32 25.2, 48 29.2
for k in ${KS:-1 3 7}; do for g in ${GROUPS_:-auto 8 16 24 32 48}; do
V=$(python -c "print($k*58.6e6*4/9)")
if [ $g = auto ]; then unset SSB_CHAIN_GROUP; else export SSB_CHAIN_GROUP=$g; fi
SSB_CHAIN_EXTRA_VAR=$V python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/sw.json 2> gpurun_out/sw.log
python -c "
import json;j=json.load(open('gpurun_out/sw.json'));s=j['stages_ms_per_step'];print('contigs in front',$k,'group','$g','step %.2f phase1 %.2f'%(j['ms_per_step'],s['ms_phase1']))"
done; done
