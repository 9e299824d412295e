mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_columns.py -m gpu -x -q > gpurun_out/pytest_cols.log 2>&1; echo "cols rc=$?"; tail -5 gpurun_out/pytest_cols.log
GRMP_FAST_PROF=1 python tools/fast_sweep.py --level 6 --steps 20 "" "GRMP_FAST_NBUF=3" "GRMP_FAST_NW=6" "GRMP_FAST_SLOT=1024" > gpurun_out/fast_prof2.log 2>&1
cat gpurun_out/fast_prof2.log | cut -c1-260
python tools/bench_configs.py --paths columns > gpurun_out/configs_v3.jsonl 2> gpurun_out/configs_v3.err; tail -3 gpurun_out/configs_v3.err
python - <<'PY'
import json
for l in open('gpurun_out/configs_v3.jsonl'):
    r=json.loads(l); print(r.get('config'), r.get('path'), r.get('ms_per_assembly'), r.get('frac_of_peak'), r.get('skipped'))
PY
GRMP_COL_NW=4 python tools/bench_configs.py --paths columns --only "tri" > gpurun_out/configs_v3_nw4.jsonl 2>/dev/null
python - <<'PY'
import json
for l in open('gpurun_out/configs_v3_nw4.jsonl'):
    r=json.loads(l); print('nw4', r.get('config'), r.get('path'), r.get('ms_per_assembly'), r.get('frac_of_peak'), r.get('skipped'))
PY
