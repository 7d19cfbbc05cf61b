for sg in 2 3 4; do
BA_LDLT_SPLIT_SEGMENTS=$sg python bench.py --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('segments $sg', d['ms_per_step'], d['e2e']['value'], d['roofline']['stages_ms']['factor'])"
done
BA_SPLIT_TIMELINE=2 python bench.py --steps 20 --warmup 5 --no-other-variant --no-parity-probe 2>&1 >/dev/null | grep "split timeline" | tail -19
