python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2j_pytest.log
python tests/soak.py --seconds 40 --seed 8 > gpurun_out/r2j_soak.log 2>&1
python tools/sweep.py --steps 30 --sizes 8192 16384 > gpurun_out/r2j_sweep_onepass.jsonl 2>&1
AMCPY_B200_LIB=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_l2p.so python tools/sweep.py --steps 30 --sizes 8192 16384 > gpurun_out/r2j_sweep_twopass.jsonl 2>&1
