N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 900 $TR --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -c 1200 gpurun_out/r02_bench_n$N.json
timeout 300 $TR --steps 200 --warmup 20 --no-extras --no-cpu-baseline --no-verify --workload strong16m > gpurun_out/r02_bench_strong_n$N.json 2>/dev/null; python -c "import sys,json; d=json.loads(open('gpurun_out/r02_bench_strong_n$N.json').readlines()[-1]); print('strong', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['other_kernels_ms'], d['e2e']['value'])"
