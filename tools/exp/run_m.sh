python tools/ab.py --steps 100 --rounds 3 base=amcpy_b200/_lib/libamcpy_b200.so unrc=amcpy_b200/_lib/exp/libamcpy_b200_unrc.so > gpurun_out/r2o_ab.log 2>&1
python tools/sweep.py --steps 30 --sizes 256 > gpurun_out/r2o_sweep_base.jsonl 2>&1
AMCPY_B200_LIB=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_wno.so python tools/sweep.py --steps 30 --sizes 256 > gpurun_out/r2o_sweep_wno.jsonl 2>&1
