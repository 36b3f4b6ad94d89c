# Round-end measurements on one B200 (run under gpurun): the driver's bench line, the panel line, one step under ncu (launch list + DRAM
# bytes), full captures of the largest kernels.  Every ncu command repeats a command that has already exited 0 without ncu.
set -u
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' || exit 1
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.log || exit 1
python bench.py --workload panel --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_panel.json 2> gpurun_out/r02_panel.log || exit 1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/plain1.json 2> gpurun_out/plain1.log || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 170 -c 200 --csv \
    --log-file gpurun_out/r02_spike_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(phase1_kernel|parse_kernel|emit_kernel|mates_kernel)$' --launch-skip 12 -c 4 \
    -o gpurun_out/r02_spike_top4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tnc > gpurun_out/ncu_f.log 2>&1
python bench.py --workload tnc --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.json 2> gpurun_out/plain2.log || exit 1
ncu --set full --clock-control none --import-source on -k regex:tnc_scan_kernel --launch-skip 9 -c 1 -o gpurun_out/r02_tnc_scan -f \
    python bench.py --workload tnc --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tn.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_spike_traffic.csv
