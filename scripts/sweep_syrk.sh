export ORCVIO_TRI_OVERLAP=0
for f in 1000 2000 4096; do
for st in 2 4; do for w in 1; do
  ORCVIO_SYRK_STAGES=$st ORCVIO_SYRK_WAVES=$w python scripts/stage_times.py --features $f --repeat 20 --flush 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('f',d['features'],'stages',$st,'waves',$w,'frame',d['us_per_frame'],'syrk',d['stages'].get('syrk'), d['stages'])"
done; done; done
