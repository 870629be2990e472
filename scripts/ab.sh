# A/B experiment runner: bash scripts/ab.sh <variant> [<variant> ...]   ("base" = the default library)
A="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BWAGPU_LIB; else export BWAGPU_LIB=$PWD/network-aware-bwa_b200/variants/libbwagpu_$v.so; fi
  echo "== $v"
  python bench.py $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g ms/step %.1f width %.1f tiers %s t2 %d stored/read %.0f own/read %.0f parity %s' % (d['value'], d['ms_per_step'], r['width_ms_per_step'], [round(x,1) for x in r['tier_ms_per_step']], d['config']['tier2_reads'], r['stored_pushes_per_read'], r['own_32B_blocks_per_read'], d['parity_sample']['mismatches']), r['pops_per_read'], r['per_read'], r['stats_pass_ms'])"
done
