# 8-GPU box, final build: the scaling series of bench.py, most valuable first (the call is cut when the GPU budget ends)
for n in 8 2 4; do
  timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 100 --warmup 5 --sustain-seconds 2 > gpurun_out/r2h_scale_$n.json 2> gpurun_out/r2h_scale_$n.err
done
