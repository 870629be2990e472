# bash scripts/ab_env.sh "ENV1=a ENV2=b" "ENV1=c" ...   (each argument = one environment to bench under)
A="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g ms/step %.1f width %.1f tiers %s t2 %d parity %s' % (d['value'], d['ms_per_step'], r['width_ms_per_step'], [round(x,1) for x in r['tier_ms_per_step']], d['config']['tier2_reads'], d['parity_sample']['mismatches']), r['stats_pass_ms'])"
done
