"""Markdown tables for DESIGN.md section 7.1 from the committed evidence files (profiles/r2_*.json)."""
import json
import os

P = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles')


def load(name):
    try:
        with open(os.path.join(P, name)) as f:
            return json.load(f)
    except Exception:
        return None


rows = [('Faster R-CNN R50, batch 8 × 600×1024, 2000 proposals, 80 classes (headline)', 'r2_bench_frcnn_r50_b8.json'),
        ('SSD-300 VGG-16, batch 32', 'r2_bench_ssd_b32.json'),
        ('Faster R-CNN R101 + block4 tail, batch 8, 300 proposals', 'r2_bench_frcnn_r101_b8_r300.json'),
        ('Faster R-CNN R50, batch 1 (latency)', 'r2_bench_frcnn_r50_latency_b1.json'),
        ('Faster R-CNN R50, batch 2', 'r2_bench_frcnn_r50_latency_b2.json')]
print('| workload (one B200) | images/s device-resident | images/s end to end | ms / step | conv_tc ms | roofline.frac (conv, ≤ 1/3) | SM MHz |')
print('|---|---|---|---|---|---|---|')
for name, f in rows:
    d = load(f)
    if not d:
        continue
    print('| %s | %.0f | %.0f | %.3f | %.3f | %.3f | %.0f |' % (
        name, d['value'], d['e2e']['value'], d['ms_per_step'], d['category_ms_per_step']['conv_tc'], d['roofline']['frac'],
        d['clocks']['sm_mhz']))
d = load('r2_bench_frcnn_r50_b8.json')
if d:
    c = d['category_ms_per_step']
    print()
    print('Headline step (kernel time per category, single-stream sums; the two-stream pipeline overlaps them into %.3f ms): '
          % d['ms_per_step'] + ', '.join('%s %.3f' % (k, v) for k, v in c.items() if v > 0) + ' ms.')
    cb = d.get('cpu_baseline')
    if cb:
        print('CPU baseline (oracle port, %d threads): %.3f images/s (%s).' % (cb['cores'], cb['value'], cb['sample']))
    r = d['roofline']
    print('Conv roofline: %.1f TFLOP/s algorithmic of %.0f (%s); DRAM traffic of the conv launches of one step %.2f GB (%s).' % (
        r['achieved'], r['peak'], r['peak_source'], (r.get('traffic') or 0) / 1e9, r.get('traffic_note', '')))
ref = load('r2_bench_reference_arm.json')
if ref:
    print('Reference arm (`bench.py --impl reference`): %s' % json.dumps({k: ref[k] for k in ref if k in ('value', 'unit', 'impl', 'ms_per_step')}))
print()
print('| GPUs (weak scaling, per-GPU batch 8; N = 1, 2, 4 on one 4-GPU box, N = 8 on an 8-GPU box) | images/s | ms / step | efficiency vs N × the N = 1 rate of the 4-GPU box |')
print('|---|---|---|---|')
base = None
for n in (1, 2, 4, 8):
    d = load('r2_scale_n%d.json' % n) or (load('r2_scale_n1_same_box_as_n2.json') if n == 1 else None)
    if not d:
        continue
    if n == 1:
        base = d['value']
    print('| %d | %.0f | %.3f | %s |' % (n, d['value'], d['ms_per_step'], ('%.3f' % (d['value'] / (n * base))) if base else '-'))
