timeout 120 python -m pytest tests/test_fused_multigpu.py -x -q -k "collective or (2gpu and zigzag)" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 100 $TR bench.py --gpus 2 --steps 2 --warmup 3 --seq 131072 > gpurun_out/n2_ours_fb_128k.json 2> gpurun_out/n2_ours.err; tail -1 gpurun_out/n2_ours.err | cut -c1-200; cat gpurun_out/n2_ours_fb_128k.json | cut -c1-260
