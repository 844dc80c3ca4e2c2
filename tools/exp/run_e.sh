python tools/sweep.py --steps 30 > gpurun_out/r2g_sweep_pdl.jsonl 2>&1
AMCPY_B200_NO_PDL=1 python tools/sweep.py --steps 30 > gpurun_out/r2g_sweep_nopdl.jsonl 2>&1
python tools/sweep.py --steps 30 --sizes 1024 2048 > gpurun_out/r2g_sweep_pdl_b.jsonl 2>&1
