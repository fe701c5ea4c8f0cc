set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 4 --steps 3 --warmup 3 --impl reference --mode fwd > gpurun_out/n4_ref_fwd.json 2> gpurun_out/n4_ref_fwd.err; tail -1 gpurun_out/n4_ref_fwd.json
$TR bench.py --gpus 4 --steps 3 --warmup 3 --mode fwd > gpurun_out/n4_ours_fwd.json 2> gpurun_out/n4_ours_fwd.err; tail -2 gpurun_out/n4_ours_fwd.err; tail -1 gpurun_out/n4_ours_fwd.json
$TR bench.py --gpus 4 --steps 2 --warmup 3 --impl reference --mode fwdbwd > gpurun_out/n4_ref_fb.json 2> gpurun_out/n4_ref_fb.err; tail -1 gpurun_out/n4_ref_fb.json
$TR bench.py --gpus 4 --steps 2 --warmup 3 --mode fwdbwd > gpurun_out/n4_ours_fb.json 2> gpurun_out/n4_ours_fb.err; tail -2 gpurun_out/n4_ours_fb.err; tail -1 gpurun_out/n4_ours_fb.json
