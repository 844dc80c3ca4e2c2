set -x
python bench.py --no-e2e --steps 20 --warmup 3 > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --no-e2e --steps 20 --warmup 3 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused16_features -s 5 -c 1 -f -o gpurun_out/r2_prof_fused16 \
    python bench.py --no-e2e --steps 8 --warmup 3 > gpurun_out/r2_ncu2.log 2>&1
python tools/profile_families.py > gpurun_out/r2_families_plain.jsonl 2> gpurun_out/r2_families_plain.err &&
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_families \
    python tools/profile_families.py > gpurun_out/r2_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/r2_families.ncu-rep > gpurun_out/r2_families_summary.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2p_pytest.log
python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python tools/sweep.py --steps 30 > gpurun_out/r2p_sweep.jsonl 2>&1
python tools/sweep.py --steps 30 --dtype c64 > gpurun_out/r2p_sweep_c64.jsonl 2>&1
