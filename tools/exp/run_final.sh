python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python bench.py --impl reference > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err
