# A/B of the chain groups on the bench workloads (same chains either way; the posterior mean is printed as a check)
for w in cfg3 cfg5; do for g in 1 2 0; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --chain-groups $g 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('$w groups', r['config']['chain_groups'], 'value %.4g ms/step %.1f kernel_s %.3f frac %.3f e2e %.4g launches %d mean %s' % (r['value'], r['ms_per_step'], r['work']['kernel_s'], r['roofline']['frac'], r['e2e']['value'], r['gpu_launches'], r['ess']['posterior_mean']))"
done; done
