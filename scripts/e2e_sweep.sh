for cfg in "BWAGPU_LANES=2" "BWAGPU_LANES=2 BWAGPU_CHUNK=1048576" "BWAGPU_LANES=3" "BWAGPU_LANES=2 BWAGPU_HOST_THREADS=4" "BWAGPU_LANES=1"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
done
