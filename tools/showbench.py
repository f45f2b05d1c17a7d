import json,sys
for f in sys.argv[1:]:
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'value %.0f (%.1f us/step) e2e %.0f (%.1f us) launches %d clocks %s'%(b['value'],1e3*b['ms_per_step'],b['e2e']['value'],1e3*b['e2e']['ms_per_step'],b['gpu_launches'],b['clocks']))
    for p in b['roofline']['phases']: print('   %-42s %7.2f us  %s' % (p['name'], p['us'], {k: v for k, v in p.items() if k not in ('name', 'us')}))
    for k, v in b.get('also', {}).items(): print('   also.%-22s %.4g %s  (%s)' % (k, v['value'], v['unit'], v.get('ms_per_step', v.get('ms'))))
