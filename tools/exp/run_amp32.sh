# EXPERIMENT: the N = 2048 kernel with float32 amplitude statistics - timing against the product, errors against the oracle
python tools/ab.py --steps 100 --rounds 2 product=amcpy_b200/_lib/libamcpy_b200.so amp32=amcpy_b200/_lib/exp/libamcpy_b200_amp32.so > gpurun_out/r2l_ab.log 2>&1
AMCPY_B200_LIB=$PWD/amcpy_b200/_lib/exp/libamcpy_b200_amp32.so python tests/amp32_experiment.py > gpurun_out/r2l_amp32_errors.jsonl 2>&1
python tests/amp32_experiment.py > gpurun_out/r2l_product_errors.jsonl 2>&1
