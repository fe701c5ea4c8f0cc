TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 4 --steps 3 --warmup 3 --mode fwdbwd > gpurun_out/n4_ours_fb2.json 2> gpurun_out/n4_ours_fb2.err; tail -2 gpurun_out/n4_ours_fb2.err; cat gpurun_out/n4_ours_fb2.json
LCA_B200_FUSED_BWD=0 $TR bench.py --gpus 4 --steps 3 --warmup 3 --mode fwdbwd > gpurun_out/n4_ours_fb_coll.json 2> gpurun_out/n4_ours_fb_coll.err; cat gpurun_out/n4_ours_fb_coll.json
