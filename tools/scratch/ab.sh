for pdl in 0 1; do for op in 0 1; do echo "FLK_PDL=$pdl FLK_SCAN_ONEPASS=$op"; FLK_PDL=$pdl FLK_SCAN_ONEPASS=$op python tools/time_configs.py shipped n4000 n60000 n250000 1m weak16m; done; done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
