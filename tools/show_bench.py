import json, sys
d = json.load(open(sys.argv[1]))
for k in ['value', 'ms_per_step', 'fwd', 'e2e', 'gpu_launches', 'cpu_baseline', 'clocks']:
    print(k, d.get(k))
print('launch', d['config'].get('launch'))
print('roofline', d.get('roofline'))
for r in d.get('kernels') or []:
    print(f"{r['kernel']:45s} {r['ms']:.4f} ms  {r['achieved']:.2f}/{r['peak']:.1f} {r['unit']} frac={r['frac']:.3f}")
