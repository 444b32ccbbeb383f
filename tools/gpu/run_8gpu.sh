# 8-GPU session: PCIe ceiling, torchrun bench (weak + in-library legs), cfg5 through the z-slab partitioner, 4-GPU legs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1; nproc; free -g | head -2
python tools/gpu/pcie_probe.py --gpus 1,2,4,8 --json gpurun_out/pcie_probe_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_8gpu.json
tail -2 gpurun_out/bench_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json'))
for k in ('value','ms_per_step','e2e','zslab','batch','batch_no_odf'): print(k, d.get(k))"
timeout 600 python bench.py --mode zslab --config cfg5 --gpus 8 2>gpurun_out/bench_cfg5_8gpu.err | tail -1 > gpurun_out/bench_cfg5_8gpu.json; tail -2 gpurun_out/bench_cfg5_8gpu.err | cut -c1-300; cut -c1-1200 gpurun_out/bench_cfg5_8gpu.json
timeout 300 python bench.py --mode batch --gpus 4 2>/dev/null | tail -1 > gpurun_out/bench_batch_4gpu.json; cut -c1-1500 gpurun_out/bench_batch_4gpu.json
