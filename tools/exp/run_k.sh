python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2m_pytest.log
python tests/soak.py --seconds 60 --seed 11 > gpurun_out/r2m_soak.log 2>&1
python tools/sweep.py --steps 30 > gpurun_out/r2m_sweep.jsonl 2>&1
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
