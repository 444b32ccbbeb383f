# 2-GPU session on the final tree: the tests that need two devices, then the torchrun bench (weak line + in-library z-slab / batch legs)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or batch or host_paths" -s 2>&1 | grep -v "^$" > gpurun_out/pytest_2gpu_r2k.log; tail -2 gpurun_out/pytest_2gpu_r2k.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu_r2k.err | tail -1 > gpurun_out/bench_2gpu_r2k.json
tail -2 gpurun_out/bench_2gpu_r2k.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_r2k.json'))
for k in ('value','ms_per_step','n_gpus','e2e','zslab','batch'): print(k, str(d.get(k))[:400])"
