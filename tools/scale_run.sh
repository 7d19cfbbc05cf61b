for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02b_scale_${n}gpu.json 2> gpurun_out/r02b_scale_${n}gpu.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02b_scale_${n}gpu.json').read().strip().splitlines()[-1]); print($n, d['ms_per_step'], d['e2e']['value'], d['variants']['ms_per_step'], d['roofline']['stages_ms']['all_reduce'], d['roofline']['stages_ms']['factor'], d['check']['energy_test'], d['parity_probe']['ok'])"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_scale_1gpu.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02b_scale_1gpu.json').read().strip().splitlines()[-1]); print(1, d['ms_per_step'], d['e2e']['value'], d['variants']['ms_per_step'], d['check']['energy_test'])"
