set -x
timeout 400 python -m pytest tests/test_fused_multigpu.py -x -q -k "8gpu" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 --impl reference > gpurun_out/n8_ref_fb.json 2> gpurun_out/n8_ref_fb.err; tail -1 gpurun_out/n8_ref_fb.json
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_ours_fb.json 2> gpurun_out/n8_ours_fb.err; tail -2 gpurun_out/n8_ours_fb.err; tail -1 gpurun_out/n8_ours_fb.json
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 --mode fwd > gpurun_out/n8_ours_fwd.json 2> gpurun_out/n8_ours_fwd.err; tail -1 gpurun_out/n8_ours_fwd.json
