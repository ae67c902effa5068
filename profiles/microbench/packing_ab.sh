# A/B of the round-level packing of the d = 3 kernel on the cfg 3 workload (same chains either way)
for p in 1 0; do for g in 1 0; do
  python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --round-packing $p --chain-groups $g 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('cfg3 packing', 'off' if $p else 'on', 'groups', r['config']['chain_groups'], 'value %.4g ms/step %.1f kernel_s %.3f frac %.3f e2e %.4g mean %s' % (r['value'], r['ms_per_step'], r['work']['kernel_s'], r['roofline']['frac'], r['e2e']['value'], r['ess']['posterior_mean']))"
done; done
