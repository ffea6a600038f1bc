python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1_v3_n2_reference.json 2> gpurun_out/bench_r1_v3_n2_reference.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_r1_v3_n2.json 2> gpurun_out/bench_r1_v3_n2.err
grep -o '"value": [0-9.]*' gpurun_out/bench_r1_v3_n2_reference.json | head -1
grep -o '"value": [0-9.]*' gpurun_out/bench_r1_v3_n2.json | head -2
grep -o '"stage_ms".*' gpurun_out/bench_r1_v3_n2.json | cut -c1-420
tail -3 gpurun_out/bench_r1_v3_n2.err gpurun_out/bench_r1_v3_n2_reference.err | grep -v OMP | grep -v "\*\*\*" | head
