# last check of the shipped library: GPU tests, smoke, one quick bench line
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1
python bench.py --quick > gpurun_out/r2j_bench_quick.json 2> gpurun_out/r2j_bench_quick.err
