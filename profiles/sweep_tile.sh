# usage: bash profiles/sweep_tile.sh "<env settings>" ...   (each arg = one bench run of the cooperative kernel)
export RXN_TILE_VERBOSE=1
CELLS=${CELLS:-600000}
WL=${WL:-hanford300a_eq}
for cfg in "$@"; do
  env $cfg timeout 200 python bench.py --steps 2 --warmup 1 --kernel 2 --cells $CELLS --workload $WL 2>&1 | python -c "
import sys,json
L=sys.stdin.read().strip().splitlines(); d=json.loads(L[-1])
print('$cfg |', L[0][11:80], '| Mcells/s %.2f kernel_ms %.1f bad %d'%(d['value']/1e6, d['roofline']['kernel_ms'], d['config']['cells_with_nonreference_flags']))"
done
