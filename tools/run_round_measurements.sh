python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02q_gputest.log; cat gpurun_out/r02q_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02q_bench_1gpu.json 2> gpurun_out/r02q_bench_1gpu.err; cut -c1-200 gpurun_out/r02q_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02q_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r02q_bench_reference.json
python bench.py --workload mixed > gpurun_out/r02q_mixed_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r02q_mixed_1gpu.json
python tools/sweep_configs.py > gpurun_out/r02q_sweep_configs.jsonl 2> gpurun_out/r02q_sweep.err
export DVBS2B200_LIB=$PWD/gr-dvbs2rx_b200/libdvbs2_b200_prof.so; for cfg in "C1_2 1 1.0 25" "C3_4 1 4.6 50" "C3_5 1 3.5 25" "C2_3 0 3.3 25" "C9_10 1 6.6 25"; do python tools/phase_profile.py $cfg 888; done > gpurun_out/r02q_phase_profile.txt 2>&1; unset DVBS2B200_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/r02q_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
for m in "pl C1_2 1 0 0 2664" "apsk C9_10 1 17.5 0 2664" "apsk C2_3 1 10.5 0 2664" "snr C1_2 1 2.0 0 2664" "snr C3_5 1 6.2 0 2664" "bb C1_2 1 0 0 8192" "ts C1_2 1 2.0 25 2664"; do python tools/run_one.py $m 2>&1 | tail -1; done > gpurun_out/r02q_other_kernels.txt
for c in "c1 C1_2 1 1.0 25 2664" "c4 C2_3 0 3.3 25 8192"; do set -- $c; ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -f -o gpurun_out/r02p_ldpc_$1 python tools/run_one.py fec $2 $3 $4 $5 $6 > /dev/null 2>&1; done
ls -la gpurun_out | tail -12
