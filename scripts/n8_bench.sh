for gb in 8 3; do
GVOM_GATHER_BLOCKS=$gb python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2992$gb bench.py --gpus 8 --steps 80 --warmup 10 > gpurun_out/bench_r1_v3_n8_gb$gb.json 2> gpurun_out/bench_r1_v3_n8_gb$gb.err
echo gather_blocks=$gb; grep -o "\"value\": [0-9.]*" gpurun_out/bench_r1_v3_n8_gb$gb.json | head -1; grep -o "\"slab_cells\".*\"slab_gather_cells\": [0-9.]*" gpurun_out/bench_r1_v3_n8_gb$gb.json
done
