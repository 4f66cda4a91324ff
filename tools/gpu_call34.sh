mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r3f_bench_2gpu.json 2> gpurun_out/r3f_bench_2gpu.err
tail -3 gpurun_out/r3f_bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r3f_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['dead_fraction']); print(json.dumps(d.get('config4'), indent=1))"
