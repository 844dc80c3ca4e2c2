python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2k_pytest.log
python tests/soak.py --seconds 40 --seed 9 > gpurun_out/r2k_soak.log 2>&1
python tools/sweep.py --steps 30 > gpurun_out/r2k_sweep.jsonl 2>&1
python tools/sweep.py --steps 30 --dtype c64 > gpurun_out/r2k_sweep_c64.jsonl 2>&1
