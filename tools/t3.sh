BA_SPLIT_TIMELINE=1 python tools/split_prof.py 2>&1 | grep "split timeline" | tail -16
echo ---- 2 ranks
BA_SPLIT_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 2 --steps 2 --warmup 3 --no-other-variant --no-parity-probe 2>&1 | grep "split timeline" | tail -16
