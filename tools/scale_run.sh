# driver-style scaling run on one 8-GPU box (light: no QRKIT leg, no oracle probe)
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-other-variant --no-parity-probe > gpurun_out/r02c_scale_${n}gpu.json 2> gpurun_out/r02c_scale_${n}gpu.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02c_scale_${n}gpu.json').read().strip().splitlines()[-1]); print($n, d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['all_reduce'], d['roofline']['stages_ms']['factor'], d['check']['energy_test'])"
done
