# A/B experiment runner for k_search: bash scripts/ab.sh <variant> [<variant> ...]   ("base" = the default library;
# a variant = network-aware-bwa_b200/variants/libbwagpu_<variant>.so made by build.build_variant(name, defines[, units])).
# Times K2 + K3 on resident C4 reads (scripts/kbench.py) and checks 10 k reads against the live reference.
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3 --check 10000"
for v in "$@"; do
  if [ "$v" = "base" ]; then unset BWAGPU_LIB; else export BWAGPU_LIB=$PWD/network-aware-bwa_b200/variants/libbwagpu_$v.so; fi
  $K --tag $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['full']
print('%-12s search %.1f ms  width %.1f ms  %.2f M reads/s  mismatches %s' % (d['tag'], f['ms_search'], f['ms_width'], f['mreads_per_s'], d.get('mismatches')))"
done
