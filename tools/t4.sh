BA_SPLIT_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('timeline', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('normal', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms'])"
