# 8-GPU box: topology record, the scaling series of bench.py (weak scaling value + config-3 strong scaling + gather +
# e2e with its copy-only ceiling at every N), the multi-GPU tests.  One line per N under gpurun_out/.
{ nvidia-smi topo -m; echo; nvidia-smi topo -p2p r 2>/dev/null | head -20; echo; lspci -tv 2>/dev/null | grep -i -B2 -A2 "nvidia\|3d controller" | head -80; echo; numactl -H 2>/dev/null; lscpu | head -25; } > gpurun_out/r2_topology.txt 2>&1
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 100 --warmup 5 --sustain-seconds 2 > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
done
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_multi_pytest.log
