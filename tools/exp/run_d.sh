export V2=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_v2.so
AMCPY_B200_LIB=$V2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_boundary.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2f_pytest_v2.log
python tools/ab.py --steps 100 --rounds 3 v1=amcpy_b200/_lib/libamcpy_b200.so v2=amcpy_b200/_lib/exp/libamcpy_b200_v2.so > gpurun_out/r2f_ab.log 2>&1
AMCPY_B200_LIB=$V2 python tests/soak.py --seconds 45 --seed 7 > gpurun_out/r2f_soak_v2.log 2>&1
python tools/sweep.py --steps 30 > gpurun_out/r2f_sweep_pdl.jsonl 2>&1
AMCPY_B200_NO_PDL=1 python tools/sweep.py --steps 30 > gpurun_out/r2f_sweep_nopdl.jsonl 2>&1
