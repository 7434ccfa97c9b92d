import json, sys
for f in sys.argv[1:]:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, '%.1f samples/s  %.2f ms  e2e %.1f  clk %s W %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))
    r=d['roofline']; print('  ', {k:round(r[k],3) for k in ('achieved','frac','tensor_pipe_frac_burst','tensor_pipe_frac_sustained','share_of_step')})
    print('  ', {k:round(v,2) for k,v in list(d['kernel_time_ms_per_step'].items())[:12]})
