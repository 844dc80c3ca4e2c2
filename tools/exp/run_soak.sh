# longer randomised soak of the final build against the oracle (three seeds in parallel processes on one GPU)
for s in ${SOAK_SEEDS:-71 72 73}; do python tests/soak.py --seconds ${SOAK_SECONDS:-170} --seed $s > gpurun_out/r2i_soak_$s.jsonl 2>&1 & done; wait
