mkdir -p gpurun_out
python tools/soft_debug.py 14 2>&1 | grep interpolation | tail -3; MAXIT=200 python tools/soft_debug.py 14 2>&1 | grep interpolation | tail -3
{ echo "== lpt on"; timeout 300 python tools/iter_stats.py 4096 | tail -1; echo "== lpt off"; timeout 300 python tools/iter_stats.py 4096 qp_lpt=0 | tail -1; } 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closed_loop_vs_oracle or population_vs_cpu_port and rocket and 4096" 2>&1 | tail -2
