python -m pytest tests -m gpu -x -q > gpurun_out/pytest_2p.log 2>&1; tail -3 gpurun_out/pytest_2p.log | cut -c1-300
for sr in 4096 32768; do echo "== production seed=$sr"; KNN_SEED_ROWS=$sr python tools/diag_stalls.py --queries 8192 --rows 20000000 --dim 512 --steps 20 2>&1 | tee gpurun_out/exp11_c5_s$sr.log; done
export KNN_PAIR_STATS=1
for sr in 4096; do echo "== C5 stats seed=$sr"; KNN_SEED_ROWS=$sr python tools/diag_stalls.py --queries 8192 --rows 20000000 --dim 512 --steps 20 2>&1 | tee gpurun_out/exp11_c5_stats_s$sr.log; done
