"""Print the headline and the per-kernel table of bench.py JSON lines (arguments: log files)."""
import json, sys
for f in sys.argv[1:]:
    print('=====', f)
    for ln in open(f).read().strip().splitlines():
        try:
            d = json.loads(ln)
        except Exception:
            print('NONJSON', ln[:300]); continue
        print('value %.4g  ms/step %.2f  e2e ms %.2f  clocks %s' % (d['value'], d['ms_per_step'], d.get('e2e', {}).get('ms_per_step', 0), d.get('clocks')))
        r = d.get('roofline')
        if r: print('roofline: launch ms %.4f frac %.4f hbm_gbs %s' % (r.get('avg_launch_ms', 0), r['frac'], r.get('hbm_gbs', r.get('achieved'))))
        for k in ('cpu_baseline', 'gpu_baseline'):
            if k in d: print(k, json.dumps(d[k])[:400])
        print('per_step_ms', d.get('per_step_ms'))
        for r in d.get('kernels', []): print('   %-34s %8.3f ms  frac %s  ceil %s' % (r['kernel'][:34], r['ms_per_step'], r.get('frac'), r.get('frac_of_3xfp16_ceiling')))
