mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02z4_gpus.log 2>&1
A="--layers 3 --size 176 144 --frames 6 --gen g1"
: > gpurun_out/r02z4_svc_layers.jsonl
for D in "0,0,0" "0,1,2" "2,1,0" "0,1,2"; do
  out=$(HLB200_SVC_DEVICES=$D oracle/_ref/hl_b200_encoder $A --out /tmp/b.264 2>gpurun_out/r02z4_svc.err | tail -1)
  echo "{\"devices\": \"$D\", \"result\": $out}" >> gpurun_out/r02z4_svc_layers.jsonl
done
cut -c1-250 gpurun_out/r02z4_svc_layers.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --no-all-inter --no-hbm-kernels > gpurun_out/r02z4_bench_4gpu.json 2> gpurun_out/r02z4_bench_4gpu.err; echo "bench 4gpu rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z4_bench_4gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e'].get('encode_fps'),'parity',d['parity_checked'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02z4_bench_ref_4gpu.json 2>> gpurun_out/r02z4_bench_4gpu.err; echo "ref arm rc=$?"; cut -c1-200 gpurun_out/r02z4_bench_ref_4gpu.json; nproc
