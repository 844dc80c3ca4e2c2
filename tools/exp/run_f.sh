python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_pytest.log
python tools/sweep.py --steps 30 > gpurun_out/r2h_sweep.jsonl 2>&1
export AMCPY_B200_LIB=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_v2.so
python bench.py --no-e2e --steps 8 --warmup 3 > gpurun_out/r2h_v2_plain.json 2> gpurun_out/r2h_v2_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:fused16x_features -s 5 -c 1 -f -o gpurun_out/r2h_prof_v2 \
    python bench.py --no-e2e --steps 8 --warmup 3 > gpurun_out/r2h_ncu_v2.log 2>&1
