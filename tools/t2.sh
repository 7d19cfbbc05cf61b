for mc in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29530+mc)) bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($mc, d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'], d['roofline']['stages_ms']['all_reduce'])"
done
BA_LDLT_SPLIT_SEGMENTS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('seg1', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'], d['roofline']['stages_ms']['all_reduce'])"
BA_LDLT_SPLIT=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nosplit', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'], d['roofline']['stages_ms']['all_reduce'])"
