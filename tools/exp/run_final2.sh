# after tools/record_profile_meta.py: the bench line with roofline.traffic and strong_scaling.matches_n1 tied to this build,
# and one ncu --set full capture per kernel family of the final build (report stays on the box, summary comes back)
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python tools/profile_families.py > gpurun_out/r2g_families_plain.jsonl 2> gpurun_out/r2g_families_plain.err &&
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2g_families \
    python tools/profile_families.py > gpurun_out/r2g_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/r2g_families.ncu-rep > gpurun_out/r2g_families_summary.txt 2>&1
