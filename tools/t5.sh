BA_SPLIT_TIMELINE=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2> gpurun_out/t5.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deferred', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'])"
grep "split timeline" gpurun_out/t5.err | sed -n 40,80p
