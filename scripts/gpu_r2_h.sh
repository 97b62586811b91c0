#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== solver + unet + vae + clip tests" | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rfs > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^SKIPPED" gpurun_out/pytest_all.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_all.log | grep -v "where\|tensor(" | cut -c1-300 | head -20 | tee -a $S
echo "== bench cfg4 (overlap / no overlap)" | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; grep -E "profiled|timed region|rror" gpurun_out/bench_cfg4.err | sed "s/^/cfg4: /" | tee -a $S
UCE_NO_OVERLAP=1 timeout 300 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --steps 10 --warmup 3 2>&1 >/dev/null | grep -E "profiled|timed region" | sed "s/^/cfg4 no-overlap: /" | tee -a $S
echo "== bench cfg2 (full line)" | tee -a $S
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "bench \+|Error|error" gpurun_out/bench.err | tail -12 | tee -a $S
echo "== unet launch list" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_profile.py > gpurun_out/unet_profile.log 2>&1
python - <<'PY' | tee -a $S
import csv,collections,re
lines=[l for l in open('gpurun_out/launches_unet.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    k=re.sub(r"\(.*","",row["Kernel Name"])[:50]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(row["Metric Value"].replace(",",""))
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:8]: print(n, round(t/n/1e3,2), k)
PY
