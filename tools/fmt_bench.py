import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('sims/s %.3e e2e %s ms/step %.2f tree_us %.1f GB/s %.0f frac %.3f share %.3f V %.2f net %s' % (d['value'], d['e2e'] and '%.3e'%d['e2e']['value'], d['ms_per_step'], r['avg_launch_us'], r['achieved'], r['frac'], r['share_of_step'], r['visits_per_sim'], d.get('roofline_net') and '%.0f TF'%d['roofline_net']['achieved']))
    else: print(line.rstrip()[-300:])
