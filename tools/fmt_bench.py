import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line); r = d['roofline']
        print(d['config']['workload'][:34], '| n=%s' % d['config'].get('n_prims'), '| Mrays/s %.1f' % d['value'], '| ms %.3f' % d['ms_per_step'],
              '| evals/s %.3e' % d['sdf_evals_per_s'], '| roofline %.2f/%.1fTF=%.3f' % (r['achieved'], r['peak'], r['frac']),
              '| e2e %.1f' % d.get('e2e', {}).get('value', -1), '| sdf/px %.1f' % d['avg_sdf_calls_per_pixel'], '| clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif line and not line.startswith('[gpurun] sending'):
        print(line)
