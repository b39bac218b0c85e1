for f in 2000 4096; do
for g in 0 1; do
  ORCVIO_GRAPH=$g python scripts/stage_times.py --features $f --repeat 30 --flush 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('f',d['features'],'graph',$g,'frame',d['us_per_frame'], 'chk', d['chk'], d['stages'])"
done; done
