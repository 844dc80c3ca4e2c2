export V=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_onepass.so
AMCPY_B200_LIB=$V python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_boundary.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2l_pytest_onepass.log
python tools/ab.py --steps 100 --rounds 3 twopass=amcpy_b200/_lib/libamcpy_b200.so onepass=amcpy_b200/_lib/exp/libamcpy_b200_onepass.so > gpurun_out/r2l_ab.log 2>&1
AMCPY_B200_LIB=$V python tests/soak.py --seconds 45 --seed 10 > gpurun_out/r2l_soak_onepass.log 2>&1
AMCPY_B200_LIB=$V python tools/sweep.py --steps 30 --sizes 512 1024 2048 4096 > gpurun_out/r2l_sweep_onepass.jsonl 2>&1
