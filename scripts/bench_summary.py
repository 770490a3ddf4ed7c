"""Print a compact table from bench.py JSON lines (development aid).  python scripts/bench_summary.py log [log...]"""
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith('{'):
            if line.startswith('real'): print("   ", line.strip())
            continue
        d = json.loads(line)
        def show(name, p):
            r = p.get('roofline') or {}
            c = p.get('cpu_baseline') or {}
            e = p.get('e2e') or {}
            print("%-16s value %.4g %s | ms/step %s | e2e %.4g | roof frac %s kernel_ms %s | cpu %.4g (%s cores) | e2e/cpu %.1f | launches %s" % (
                name, p['value'], p.get('unit'), ('%.3f' % p['ms_per_step']) if p.get('ms_per_step') else None, e.get('value', float('nan')),
                ('%.4f' % r['frac']) if r.get('frac') is not None else None, ('%.3f' % r['kernel_ms']) if r.get('kernel_ms') else None,
                c.get('value', float('nan')), c.get('cores'), (e.get('value', 0) / c['value']) if c.get('value') else float('nan'), p.get('gpu_launches')))
        print(f, d.get('impl', 'b200'), 'n_gpus', d.get('n_gpus'), 'clocks', d.get('clocks'))
        show(d['metric'].split()[0] + '*', d)
        for k, p in d.get('parts', {}).items(): show(k, p)
        r = d.get('roofline') or {}
        print("    fp64 peak %s TF/s, copy %s GB/s" % (r.get('fp64_dfma_peak_tflops_measured_in_run'), r.get('copy_gbs_measured_in_run')))
