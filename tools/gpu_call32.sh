mkdir -p gpurun_out
timeout 600 python tools/tol_loop.py 4096 > gpurun_out/r3f_tol_loop.txt 2>&1; cat gpurun_out/r3f_tol_loop.txt | tail -12
