export KNN_PAIR_STATS=1 KNN_BF16_NO_TS=1
for dbg in 0 1 2; do echo "== C5 SS debug=$dbg"; KNN_PAIR_DEBUG=$dbg python tools/diag_stalls.py --queries 8192 --rows 20000000 --dim 512 --steps 20 2>&1 | tee gpurun_out/exp7_c5_ss_d$dbg.log; done
