mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench1.json 2> gpurun_out/r2k_bench1.err
tail -3 gpurun_out/r2k_bench1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
print(d.get('om_step_cfg3')); print(d.get('bank_order')); print(d.get('cfg5_single_gpu'))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | head -c 600
