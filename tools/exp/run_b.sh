python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2c_pytest.log
python tests/soak.py --seconds 90 --seed 5 > gpurun_out/r2c_soak.log 2>&1
python tools/ab.py --steps 100 --rounds 3 pdl=amcpy_b200/_lib/libamcpy_b200.so > gpurun_out/r2c_ab.log 2>&1
AMCPY_B200_NO_PDL=1 python tools/ab.py --steps 100 --rounds 3 nopdl=amcpy_b200/_lib/libamcpy_b200.so >> gpurun_out/r2c_ab.log 2>&1
