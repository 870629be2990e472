for cfg in "BWAGPU_LANES=3" "BWAGPU_LANES=4" "BWAGPU_LANES=3 BWAGPU_CHUNK=1048576" "BWAGPU_LANES=4 BWAGPU_CHUNK=1048576" "BWAGPU_LANES=2"; do
  echo "== $cfg"
  env $cfg python scripts/e2e_breakdown.py 10000000 2>&1 | grep -E "^flat|^struct" | tail -2 | cut -c1-60
done
