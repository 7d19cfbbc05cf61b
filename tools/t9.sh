for e in 0 32; do
BA_SPLIT_EXP=$e python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+e)) bench.py --gpus 2 --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 exp $e', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'])"
done
python bench.py --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'])"
timeout 300 python -m pytest tests/test_gpu_band.py -x -q -k "split" 2>&1 | tail -2
