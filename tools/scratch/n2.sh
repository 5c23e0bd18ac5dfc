timeout 900 python -m pytest tests/test_gpu_strips.py tests/test_gpu_nccl.py -x -q -m gpu > gpurun_out/n2_pytest.log 2>&1; tail -n 2 gpurun_out/n2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 300 $TR --steps 200 --warmup 20 --no-extras --no-cpu-baseline --no-verify --birds-per-gpu 2000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('2M/GPU', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['other_kernels_ms'])"
timeout 900 $TR --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 1500 gpurun_out/r02_bench_n2.json
