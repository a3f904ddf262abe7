#!/bin/bash
# Round-end evidence run (one B200): full GPU test suite, smoke, bench lines of the three workloads + latency, the
# reference arm, one-step `ncu --set full` summary of the headline workload (feeds roofline.traffic), ncu launch list.
# Usage:  scripts/gpurun_retry.sh 2400 'bash scripts/gpu_evidence.sh <tag>'     (tag: e.g. r2)
TAG=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out/ev
git_commit=$(cat .evidence_commit 2>/dev/null || echo unknown)
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 --timeout-method=thread > ${O}_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > ${O}_summary.txt
cp gpurun_out/parity_report.json ${O}_parity_report.json 2>/dev/null; cp gpurun_out/parity_report_baseline.json ${O}_parity_report_baseline.json 2>/dev/null
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
echo "smoke exit $?" >> ${O}_summary.txt
# one step under ncu --set full (single stream, whole batch: the configuration of bench.py's event profile)
timeout -s KILL 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/ncu_step python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > ${O}_ncu_step.log 2>&1
echo "ncu step exit $?" >> ${O}_summary.txt
ncu -i /tmp/ncu_step.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > ${O}_ncu_step_raw.csv.gz
python scripts/ncu_step_summary.py /tmp/ncu_step.ncu-rep frcnn_r50 ${O}_ncu_step > ${O}_ncu_step_summary.txt 2>&1
echo "summary exit $?" >> ${O}_summary.txt
# merge the fresh frcnn_r50 entry into profiles/<tag>_ncu_step_summary.json BEFORE the bench reads roofline.traffic from it
python - "$TAG" "$git_commit" <<'PY'
import json, sys, datetime
tag, commit = sys.argv[1], sys.argv[2]
path = 'profiles/%s_ncu_step_summary.json' % tag
try: d = json.load(open(path))
except Exception: d = {}
new = json.load(open('gpurun_out/ev_ncu_step_summary.json'))
d.update(new)
cap = d.get('_capture', {})
cap.update({'commit': commit, 'date': datetime.date.today().isoformat(), 'frcnn_r50': 'scripts/gpu_evidence.sh (this run)'})
d['_capture'] = cap
json.dump(d, open(path, 'w'), indent=1, sort_keys=True)
json.dump(d, open('gpurun_out/ev_ncu_step_summary_merged.json', 'w'), indent=1, sort_keys=True)
PY
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 --layers > ${O}_bench_r50.json 2> ${O}_bench_r50.err
echo "bench r50 exit $? (normal interpreter exit)" >> ${O}_summary.txt
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 0 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
echo "bench ref exit $?" >> ${O}_summary.txt
timeout -s KILL 400 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > ${O}_bench_ssd.json 2>/dev/null
timeout -s KILL 400 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --layers --no-cpu-baseline > ${O}_bench_r101.json 2>/dev/null
for b in 1 2; do timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > ${O}_bench_r50_latency_b$b.json 2>/dev/null; done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file ${O}_launches_r50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_ncu_bench.log 2>&1
echo "ncu launches exit $?" >> ${O}_summary.txt
tail -n 6 ${O}_pytest_gpu.log; tail -n 2 ${O}_smoke.log
python - <<'PY'
import json
for wl in ('r50','ssd','r101','r50_latency_b1','r50_latency_b2'):
    try:
        d=json.load(open('gpurun_out/ev_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4), d['roofline'].get('traffic'), d.get('cpu_baseline'))
    except Exception as e: print(wl,'ERR',e)
PY
cat ${O}_ncu_step_summary.txt | head -24; cat ${O}_summary.txt; du -sm gpurun_out
