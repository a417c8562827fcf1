mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-hbm-kernels --no-cpu-baseline > gpurun_out/r02z5_bench_4gpu.json 2> gpurun_out/r02z5_bench_4gpu.err; echo "bench 4gpu rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z5_bench_4gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity_checked']); print(d.get('svc_layers_over_gpus'))"
tail -c 300 gpurun_out/r02z5_bench_4gpu.err
