python -m pytest tests/test_multi_gpu.py -q -m gpu -k "nccl_two" -x 2>&1 | tail -6
for m in pull p2p; do
  GVOM_MULTI=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_r1_n2_$m.json 2> gpurun_out/bench_r1_n2_$m.err
  tail -c 1500 gpurun_out/bench_r1_n2_$m.json | grep -o '"stage_ms".*' | cut -c1-400
  grep -o '"value": [0-9.]*' gpurun_out/bench_r1_n2_$m.json | head -2
done
