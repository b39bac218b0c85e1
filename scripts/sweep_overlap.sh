for f in 2000 4096; do
for ov in 0 1; do
  ORCVIO_TRI_OVERLAP=$ov python scripts/stage_times.py --features $f --repeat 20 --flush 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('f',d['features'],'overlap',$ov,'frame',d['us_per_frame'], d['stages'])"
done; done
