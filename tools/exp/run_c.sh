python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2d_pytest.log
( time python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2> gpurun_out/r2d_bench.time
