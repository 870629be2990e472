A="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
for cfg in "BWAGPU_T1_CAP=2048" "BWAGPU_T1_CAP=512" "BWAGPU_T1_CAP=1024 BWAGPU_T1_BLOCKS_PER_SM=4"; do
  echo "== $cfg"
  env $cfg python bench.py $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3g ms/step %.1f tiers %s t2 %d t3 %d stored/read %.0f pushes/read %.0f pops/read %.0f own/read %.0f parity %s' % (d['value'], d['ms_per_step'], [round(x,1) for x in r['tier_ms_per_step']], d['config']['tier2_reads'], d['config']['tier3_reads'], r['stored_pushes_per_read'], r['pushes_per_read'], r['pops_per_read'], r['own_32B_blocks_per_read'], d['parity_sample']))"
done
