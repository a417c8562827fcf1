mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02z_gpus.log 2>&1
python -m pytest tests/test_svc_bl_resample.py -q -m gpu -k "another_context" > gpurun_out/r02z_tests_handoff.log 2>&1; echo "handoff test rc=$?"; tail -2 gpurun_out/r02z_tests_handoff.log
A="--layers 3 --size 176 144 --frames 6 --gen g1"
oracle/_ref/hl_ref_driver $A --out /tmp/r.264 2>/dev/null | tail -1 > gpurun_out/r02z_svc_ref.json
: > gpurun_out/r02z_svc_layers.jsonl
for D in "" "0,0,0" "0,0,1" "0,1,1" "0,1,0"; do
  for H in "" 1; do
    if [ -n "$H" ] && [ "$D" != "" ] && [ "$D" != "0,1,1" ]; then continue; fi
    out=$(HLB200_SVC_DEVICES=$D HLB200_SVC_HOST_HANDOFF=$H oracle/_ref/hl_b200_encoder $A --out /tmp/b.264 2>gpurun_out/r02z_svc.err | tail -1)
    echo "{\"devices\": \"$D\", \"host_handoff\": \"$H\", \"result\": $out}" >> gpurun_out/r02z_svc_layers.jsonl
  done
done
cat gpurun_out/r02z_svc_ref.json | cut -c1-200; cut -c1-260 gpurun_out/r02z_svc_layers.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-all-inter --no-hbm-kernels > gpurun_out/r02z_bench_2gpu.json 2> gpurun_out/r02z_bench_2gpu.err; echo "bench 2gpu rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e'].get('encode_fps'),'parity',d['parity_checked'])"
