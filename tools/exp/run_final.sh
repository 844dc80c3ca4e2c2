# Driver-equivalent verification of the final build on one B200 + the artefacts committed under profiles/.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python bench.py --impl reference > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
python tools/sweep.py --steps 30 > gpurun_out/r2f_sweep.jsonl 2>&1
python tools/sweep.py --steps 30 --dtype c64 > gpurun_out/r2f_sweep_c64.jsonl 2>&1
{ for f in "2 4 6 8 12 14" "4 6 7 8 10 11 12 13 14 15 16 17 18" "10 11 12 13 14 15 16 17 18" "1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18"; do python tools/sweep.py --steps 30 --sizes 256 2048 --features $f 2>&1 | tail -2; done; } > gpurun_out/r2f_feature_profiles.jsonl
python tests/soak.py --seconds 100 --seed 61 > gpurun_out/r2f_soak.jsonl 2>&1
python tools/error_report.py > gpurun_out/r2f_error_report.txt 2>&1
python bench.py --no-e2e --steps 20 --warmup 3 > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --no-e2e --steps 20 --warmup 3 > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused16_features -s 5 -c 1 -f -o gpurun_out/r2f_prof_fused16 \
    python bench.py --no-e2e --steps 8 --warmup 3 > gpurun_out/r2f_ncu2.log 2>&1
