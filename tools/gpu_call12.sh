mkdir -p gpurun_out
for B in 1 148 592 1184 2368; do for K in 4 5; do timeout 120 python tools/qp_phase_profile.py $B qp_kernel=$K | awk -v b=$B -v k=$K 'NR==1{print "B="b" kernel="k" "$0} /A: W|C: Chol|backward chain|forward chain|update|ratio/{print}'; done; done > gpurun_out/r2g_phases.log 2>&1
cat gpurun_out/r2g_phases.log
