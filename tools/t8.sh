BA_SPLIT_TIMELINE=3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2> gpurun_out/t8.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('s2only', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'])"
grep "split timeline" gpurun_out/t8.err | sed -n 30,60p
