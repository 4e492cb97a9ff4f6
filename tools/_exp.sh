python -m pytest tests -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; tail -12 gpurun_out/pytest_m.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_v4.log 2>&1; tail -1 gpurun_out/bench_c5_v4.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print(j['value'], j['ms_per_step'], r['achieved'], r['kernel_ms'], r['merge_kernel_ms'], j['clocks'], j['e2e'])"
for q in 1 64; do for w in 1 2; do echo "C4 q=$q waves=$w"; KNN_WAVES=$w python tools/diag_stalls.py --queries $q --rows 10000000 --dim 768 --steps 20 2>&1 | tee gpurun_out/exp12_c4_q${q}_w$w.log; done; done
