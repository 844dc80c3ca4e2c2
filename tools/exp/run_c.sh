python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2d_pytest.log
python tools/ab.py --steps 100 --rounds 3 ser=amcpy_b200/_lib/exp/libamcpy_b200_ser.so ilv=amcpy_b200/_lib/libamcpy_b200.so > gpurun_out/r2d_ab.log 2>&1
( time python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2> gpurun_out/r2d_bench.time
python tests/soak.py --seconds 45 --seed 6 > gpurun_out/r2d_soak.log 2>&1
