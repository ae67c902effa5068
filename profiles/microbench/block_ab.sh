for w in cfg3 cfg5; do for b in 0 64 96; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --block-threads $b 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('$w block $b groups', r['config']['chain_groups'], 'value %.4g ms/step %.1f kernel_s %.3f frac %.3f' % (r['value'], r['ms_per_step'], r['work']['kernel_s'], r['roofline']['frac']))"
done; done
