for f in "2 4 6 8 12 14" "4 6 7 8 10 11 12 13 14 15 16 17 18" "10 11 12 13 14 15 16 17 18"; do python tools/sweep.py --steps 30 --sizes 256 2048 --features $f >> gpurun_out/r2_feature_profiles.jsonl 2>&1; done
python tools/sweep.py --steps 30 --sizes 256 2048 >> gpurun_out/r2_feature_profiles.jsonl 2>&1
