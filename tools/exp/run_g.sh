python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python bench.py --impl reference > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err
