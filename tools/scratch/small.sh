for t in 0 16384; do echo "FLK_REBIN_SMALL=$t"; FLK_REBIN_SMALL=$t python tools/time_configs.py shipped n2000 n4000 n8000 n12000 n16000; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -x -q -m gpu 2>&1 | tail -n 3
